/* translation unit: LaunchMsm kernels for SM2P256V1 (split so that the kernel groups compile in parallel) */
#define ECC_TU_MSM
#include "msm.cuh"
namespace eccb200 {
template struct LaunchMsm<Curve_SM2P256V1>;
}
