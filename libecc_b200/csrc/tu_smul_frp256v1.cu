/* translation unit: LaunchSmul kernels for FRP256V1 (split so that the kernel groups compile in parallel) */
#define ECC_TU_SMUL
#include "kernels.cuh"
namespace eccb200 {
template struct LaunchSmul<Curve_FRP256V1>;
}
