/* translation unit: LaunchVar kernels (K2 + direct table build) for SM2P256V1; out-of-line multiplier */
#define ECC_TU_VAR
#include "kernels.cuh"
namespace eccb200 {
template struct LaunchVar<Curve_SM2P256V1>;
}
