/* translation unit: LaunchMisc kernels for SM2P256V1 (split so that the kernel groups compile in parallel) */
#define ECC_TU_MISC
#include "kernels.cuh"
namespace eccb200 {
template struct LaunchMisc<Curve_SM2P256V1>;
}
