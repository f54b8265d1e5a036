/* translation unit: LaunchVerify kernels for SECP224R1 (split so that the kernel groups compile in parallel) */
#define ECC_TU_VERIFY
#include "kernels.cuh"
namespace eccb200 {
template struct LaunchVerify<Curve_SECP224R1>;
}
