/* translation unit: LaunchFixed kernels (K1 + wide-table merge) for SECP384R1; multiplier inlined (see kernels.cuh) */
#define ECC_TU_FIXED
#define ECC_INLINE_MUL
/* Round 2: 6 of the 8 products of the loop body call an out-of-line copy (the 384-bit product is ~470 instructions, eight
 * inlined copies are ~60 KB): measured 0 / 2 / 4 / 6 / 8 -> 177 / 181 / 188 / 194 / 192 M/s (2^20 scalars, w = 24) */
#define ECC_K1_OOL_MULS 6
#include "kernels.cuh"
namespace eccb200 {
template struct LaunchFixed<Curve_SECP384R1>;
}
