/* translation unit: LaunchFixed kernels (K1 + wide-table merge) for BRAINPOOLP512R1; multiplier inlined (see kernels.cuh) */
#define ECC_TU_FIXED
#define ECC_INLINE_MUL
#include "kernels.cuh"
namespace eccb200 {
template struct LaunchFixed<Curve_BRAINPOOLP512R1>;
}
