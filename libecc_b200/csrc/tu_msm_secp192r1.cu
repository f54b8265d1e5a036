/* translation unit: LaunchMsm kernels for SECP192R1 (split so that the kernel groups compile in parallel) */
#define ECC_TU_MSM
#include "msm.cuh"
namespace eccb200 {
template struct LaunchMsm<Curve_SECP192R1>;
}
