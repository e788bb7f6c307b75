// nddwt_fused.cu -- fused per-level kernels (placeholder until the first fused kernels land).
#include "nddwt_plan.h"
namespace nddwt {
int fused_dec_level(nddwt_plan *, int, const void *, const LevelIO &, void *const *, cudaStream_t) { return 1; }
int fused_rec_level(nddwt_plan *, int, const void *const *, void *, cudaStream_t) { return 1; }
}
