/* translation unit: LaunchVar kernels (K2 + direct table build) for BRAINPOOLP512R1; out-of-line multiplier */
#define ECC_TU_VAR
#include "kernels.cuh"
namespace eccb200 {
template struct LaunchVar<Curve_BRAINPOOLP512R1>;
}
