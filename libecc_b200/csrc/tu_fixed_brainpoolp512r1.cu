/* translation unit: LaunchFixed kernels (K1 + wide-table merge) for BRAINPOOLP512R1; multiplier inlined (see kernels.cuh) */
#define ECC_TU_FIXED
#define ECC_INLINE_MUL
/* 16- / 18-word fields: every product of the loop body is a call (an inlined body would be > 100 KB); by analogy
 * with the 384-bit measurement, not measured separately */
#define ECC_K1_OOL_MULS 8
#include "kernels.cuh"
namespace eccb200 {
template struct LaunchFixed<Curve_BRAINPOOLP512R1>;
}
