/* translation unit: LaunchVerify kernels for SM2P256V1 (split so that the kernel groups compile in parallel) */
#define ECC_TU_VERIFY
#include "kernels.cuh"
namespace eccb200 {
template struct LaunchVerify<Curve_SM2P256V1>;
}
