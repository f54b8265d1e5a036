/* translation unit: LaunchSmul kernels for SECP256R1 (split so that the kernel groups compile in parallel) */
#define ECC_TU_SMUL
#include "kernels.cuh"
namespace eccb200 {
template struct LaunchSmul<Curve_SECP256R1>;
}
