/* translation unit: LaunchMisc kernels for SECP224R1 (split so that the kernel groups compile in parallel) */
#define ECC_TU_MISC
#include "kernels.cuh"
namespace eccb200 {
template struct LaunchMisc<Curve_SECP224R1>;
}
