/* translation unit: LaunchFixed kernels (K1 + wide-table merge) for BRAINPOOLP384R1; multiplier inlined (see kernels.cuh) */
#define ECC_TU_FIXED
#define ECC_INLINE_MUL
/* as for secp384r1 (same word count; not measured separately) */
#define ECC_K1_OOL_MULS 6
#include "kernels.cuh"
namespace eccb200 {
template struct LaunchFixed<Curve_BRAINPOOLP384R1>;
}
