/* translation unit: LaunchFixed kernels (K1 + wide-table merge) for SM2P256V1; multiplier inlined (see kernels.cuh) */
#define ECC_TU_FIXED
#define ECC_INLINE_MUL
#include "kernels.cuh"
namespace eccb200 {
template struct LaunchFixed<Curve_SM2P256V1>;
}
