/* translation unit: LaunchFixed kernels (K1 + wide-table merge) for SECP256R1; multiplier inlined (see kernels.cuh) */
#define ECC_TU_FIXED
#define ECC_INLINE_MUL
/* P-256 only: the squaring stays out of line.  The fully inlined loop body (8 products + 3 squarings) is 42 KB of SASS,
 * past the 32 KB instruction cache (no_instruction stalls in profiles/); with the squaring called it is 35 KB and K1
 * runs 2.8 % faster (478.6 -> 492.0 M/s).  Measured neutral for FRP256V1 and 4 % slower for P-384, hence not general. */
#define ECC_NOINLINE_SQR
#include "kernels.cuh"
namespace eccb200 {
template struct LaunchFixed<Curve_SECP256R1>;
}
