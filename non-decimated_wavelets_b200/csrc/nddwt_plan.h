// nddwt_plan.h -- internal plan object of libnddwt_b200 (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include "nddwt_common.cuh"

struct nddwt_plan {
    int ndims = 0;
    int64_t dims[NDDWT_MAX_DIMS] = {1, 1, 1, 1};
    int L[NDDWT_MAX_DIMS] = {0, 0, 0, 0};
    int dtype = NDDWT_C64;
    int pres_l2 = 0;
    int device = 0;
    int64_t numel = 0;     // prod(dims)
    size_t esize = 0;      // bytes per element

    // unscaled wave_filters output per dim
    double lo[NDDWT_MAX_DIMS][NDDWT_MAXL];
    double hi[NDDWT_MAX_DIMS][NDDWT_MAXL];
    // per-dim taps with the level scale folded in: analysis x 2^(-1/2) per dim when pres_l2_norm,
    // synthesis x 2^(-1/2) (pres_l2_norm) or x 1/2 (otherwise) per dim  (nd_dwt_2D.m:230-232,295-299)
    nddwt::AllTaps<double> dec_d, rec_d;
    nddwt::AllTaps<float> dec_f, rec_f;

    int dil[NDDWT_MAX_LEVELS];
    int kernel_mode = 0;   // 0 auto, 1 generic only
    int last_path = 0;     // 1 fused, 0 generic
    int64_t launches = 0;

    // device scratch owned by the plan (allocated on first use, reused across calls)
    void *approx[2] = {nullptr, nullptr};      // ping-pong intermediate approximation bands
    void *gen_scratch = nullptr;               // generic path temporaries, 2*(d-1) arrays
    void *fused_scratch = nullptr;             // fused 4-D path intermediates
    size_t fused_scratch_bytes = 0;
    void *host_x = nullptr;                    // device staging for the *_host entry points
    void *host_c = nullptr;
    size_t host_c_bytes = 0;
    cudaStream_t host_stream = nullptr;
};

namespace nddwt {

void set_error(const std::string &msg);
int cuda_fail(cudaError_t e, const char *what);

#define NDDWT_CUDA(call)                                                          \
    do {                                                                          \
        cudaError_t e__ = (call);                                                 \
        if (e__ != cudaSuccess) return nddwt::cuda_fail(e__, #call);              \
    } while (0)

// geometry of one analysis/synthesis level on a slab (or the whole array)
struct LevelIO {
    const void *halo_lo = nullptr;   // planes below the slab along the last dim (nullptr: periodic)
    const void *halo_hi = nullptr;   // planes above
};

// ---- generic separable kernels (nddwt_generic.cu) ----
int generic_dec_level(nddwt_plan *p, int dil, const void *a_in, const LevelIO &io,
                      void *const *out_bands, cudaStream_t s);
int generic_rec_level(nddwt_plan *p, int dil, const void *const *in_bands, void *a_out, cudaStream_t s);
int generic_rec_stage1(nddwt_plan *p, int dil, const void *const *in_bands, void *u_lo, void *u_hi,
                       cudaStream_t s);
int generic_rec_stage2(nddwt_plan *p, int dil, const void *u_lo, const void *u_hi, const LevelIO &io,
                       void *a_out, cudaStream_t s);

// ---- fused kernels (nddwt_fused.cu); return 1 if the case has no fused instantiation ----
int fused_dec_level(nddwt_plan *p, int dil, const void *a_in, const LevelIO &io,
                    void *const *out_bands, cudaStream_t s);
int fused_rec_level(nddwt_plan *p, int dil, const void *const *in_bands, void *a_out, cudaStream_t s);
int fused_rec_stage1(nddwt_plan *p, int dil, const void *const *in_bands, void *u_lo, void *u_hi, cudaStream_t s);
int fused_rec_stage2(nddwt_plan *p, int dil, const void *u_lo, const void *u_hi, const LevelIO &io, void *a_out,
                     cudaStream_t s);

int ensure_scratch(nddwt_plan *p);

}  // namespace nddwt
