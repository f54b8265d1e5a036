/* translation unit: LaunchFixed kernels (K1 + wide-table merge) for SECP256R1; multiplier inlined (see kernels.cuh) */
#define ECC_TU_FIXED
#define ECC_INLINE_MUL
/* Round 2: of the 8 products of the loop body (extended-Jacobian mixed addition) the first 2 call an out-of-line copy:
 * measured 0 / 1 / 2 / 3 / 4 / 6 / 8 -> 646 / 652 / 668 / 661 / 660 / 652 / 642 M/s (2^20 scalars, w = 26): with everything inlined
 * the loop body overflows the instruction cache (ncu: no_instruction 1.02 stalls per issue), with everything called the
 * call overhead wins */
#define ECC_K1_OOL_MULS 2
/* P-256 only: the squaring stays out of line.  The fully inlined loop body (8 products + 3 squarings) is 42 KB of SASS,
 * past the 32 KB instruction cache (no_instruction stalls in profiles/); with the squaring called it is 35 KB and K1
 * runs 2.8 % faster (478.6 -> 492.0 M/s).  Measured neutral for FRP256V1 and 4 % slower for P-384, hence not general. */
#define ECC_NOINLINE_SQR
#include "kernels.cuh"
namespace eccb200 {
template struct LaunchFixed<Curve_SECP256R1>;
}
