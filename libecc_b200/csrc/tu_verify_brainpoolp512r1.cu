/* translation unit: LaunchVerify kernels for BRAINPOOLP512R1 (split so that the kernel groups compile in parallel) */
#define ECC_TU_VERIFY
#include "kernels.cuh"
namespace eccb200 {
template struct LaunchVerify<Curve_BRAINPOOLP512R1>;
}
