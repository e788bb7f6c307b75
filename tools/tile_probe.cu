// tile_probe.cu -- what can the tile kernels' MEMORY ACCESS PATTERN reach on its own?
// Two copy-only kernels with the geometry of k_rec3_bulk / k_dec3_fused on cfg5 (192x192x64 planes,
// 96 batch hyperplanes, 8-byte elements, 7-element halos) and no arithmetic:
//   probe_r: per plane, stage 8 haloed subband tiles (T1+8) x (T2+7) with cp.async.bulk row copies
//            + mbarrier (NBUF-deep ring), consume them with one shared-memory pass, write one T1 x T2 tile;
//   probe_w: per plane, read one haloed tile with ld.global.nc, write 8 T1 x T2 subband tiles (16-byte stores).
// Reported: ms and GB/s on the compulsory 9*N*8 bytes -- the ceiling a perfect compute pipeline would
// inherit.  Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/tile_probe.cu -o tools/tile_probe
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

using T = float2;
constexpr int H = 7, HB = 4;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n.reg .pred p;\nWAIT_LOOP:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE;\nbra WAIT_LOOP;\nDONE:\n}\n" ::"r"(
            smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ int wrapi(int m, int n) { m %= n; return m < 0 ? m + n : m; }

struct Prm {
    const T *in[8];
    T *out[8];
    int n1, n2, n3, nbatch;
    int64_t s3, s4;
    int tiles1, tiles2;
};

// ---- read-pattern probe --------------------------------------------------------------------
template <int T1, int T2, int NT, int NBUF, int MINB, int CL = 1>
__global__ void __launch_bounds__(NT, MINB) probe_r(const Prm p)
{
    constexpr int W1S = (T1 >= 192) ? T1 : T1 + 8, W2 = T2 + H;   // full rows need no dim-1 halo
    constexpr int BP = (W2 * W1S * 8 + 127) / 128 * 128 / 8;
    constexpr int NROWS = 8 * W2, KR = (NROWS + NT - 1) / NT;
    constexpr uint32_t PLANE_BYTES = 8u * W2 * W1S * 8u;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    T *RAW = reinterpret_cast<T *>(smem_raw);               // [NBUF][8][BP]
    uint64_t *bar = reinterpret_cast<uint64_t *>(RAW + (size_t)NBUF * 8 * BP);
    const int tid = threadIdx.x;
    int bid = blockIdx.x;
    const int t1 = bid % p.tiles1; bid /= p.tiles1;
    const int t2 = bid % p.tiles2; bid /= p.tiles2;
    const int batch = bid;
    const int a1 = t1 * T1, a2 = t2 * T2;
    const int n1 = p.n1, n2 = p.n2, n3 = p.n3;
    const int64_t boff = (int64_t)batch * p.s4;
    const int gc0 = (T1 >= 192) ? 0 : wrapi(a1 - HB, n1);
    const int len0 = min(W1S, n1 - gc0);
    const T *r_src[KR];
    int r_dst[KR];
#pragma unroll
    for (int k = 0; k < KR; ++k) {
        const int it = tid + k * NT;
        const int r = it % W2, b = (it < NROWS) ? it / W2 : 0;
        r_src[k] = p.in[b] + boff + (int64_t)wrapi(a2 - HB + r, n2) * n1;
        r_dst[k] = (it < NROWS) ? b * BP + r * W1S : -1;
    }
    if (tid == 0) {
        for (int i = 0; i < NBUF; ++i) mbar_init(bar + i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int z, int buf) {
        if (tid == 0) mbar_expect_tx(bar + buf, PLANE_BYTES);
        const int64_t zoff = (int64_t)wrapi(z - HB, n3) * p.s3;
#pragma unroll
        for (int k = 0; k < KR; ++k)
            if (r_dst[k] >= 0) {
                T *dst = RAW + (size_t)buf * 8 * BP + r_dst[k];
                const T *src = r_src[k] + zoff;
                bulk_g2s(dst, src + gc0, (uint32_t)(len0 * 8), bar + buf);
                if (len0 < W1S) bulk_g2s(dst + len0, src, (uint32_t)((W1S - len0) * 8), bar + buf);
            }
    };
    for (int i = 0; i < NBUF; ++i) issue(i, i);
    // output mapping: 16-byte chunks, lanes along dim 1
    constexpr int CPR = T1 / 2, NOUT = CPR * T2, KO = (NOUT + NT - 1) / NT;
    uint32_t parity = 0;
    for (int z = 0; z < n3; ++z) {
        const int buf = z % NBUF;
        mbar_wait(bar + buf, parity);
        if (buf == NBUF - 1) parity ^= 1;
        if (CL > 1) asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");   // lockstep: the CL dim-1 neighbours request the same rows together
        const T *R = RAW + (size_t)buf * 8 * BP;
        float4 o[KO];
#pragma unroll
        for (int k = 0; k < KO; ++k) {
            const int it = tid + k * NT;
            const int cp = it % CPR, j = (it / CPR) % T2;
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                const float4 v = *reinterpret_cast<const float4 *>(R + b * BP + (j + 3) * W1S + 4 + cp * 2);
                a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
            }
            o[k] = a;
        }
        __syncthreads();                           // staged tile consumed
        if (z + NBUF < n3) issue(z + NBUF, buf);
#pragma unroll
        for (int k = 0; k < KO; ++k) {
            const int it = tid + k * NT;
            const int cp = it % CPR, j = it / CPR;
            if (it < NOUT && a1 + cp * 2 < n1 && a2 + j < n2)
                __stcs(reinterpret_cast<float4 *>(p.out[0] + boff + (int64_t)z * p.s3 + (int64_t)(a2 + j) * n1 + a1 + cp * 2), o[k]);
        }
        if (CL > 1) asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
    }
}

// ---- write-pattern probe -------------------------------------------------------------------
template <int T1, int T2, int NT, int MINB, int STMODE>
__global__ void __launch_bounds__(NT, MINB) probe_w(const Prm p)
{
    const int tid = threadIdx.x;
    int bid = blockIdx.x;
    const int t1 = bid % p.tiles1; bid /= p.tiles1;
    const int t2 = bid % p.tiles2; bid /= p.tiles2;
    const int batch = bid;
    const int a1 = t1 * T1, a2 = t2 * T2;
    const int n1 = p.n1, n2 = p.n2, n3 = p.n3;
    const int64_t boff = (int64_t)batch * p.s4;
    constexpr int W1 = T1 + H, W2 = T2 + H, NPOS = W1 * W2, PPT = (NPOS + NT - 1) / NT;
    constexpr int CPR = T1 / 2, NOUT = CPR * T2, KO = (NOUT + NT - 1) / NT;
    int g_off[PPT];
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
        const int q = tid + k * NT;
        const int r = (q / W1) % W2, c = q % W1;
        g_off[k] = wrapi(a2 - 3 + r, n2) * n1 + wrapi(a1 - 3 + c, n1);
    }
    for (int z = 0; z < n3; ++z) {
        const T *pl = p.in[0] + boff + (int64_t)z * p.s3;
        float2 s = make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < PPT; ++k)
            if (tid + k * NT < NPOS) { const float2 v = __ldg(pl + g_off[k]); s.x += v.x; s.y += v.y; }
#pragma unroll
        for (int k = 0; k < KO; ++k) {
            const int it = tid + k * NT;
            const int cp = it % CPR, j = it / CPR;
            if (it < NOUT && a1 + cp * 2 < n1 && a2 + j < n2) {
                const int64_t off = boff + (int64_t)z * p.s3 + (int64_t)(a2 + j) * n1 + a1 + cp * 2;
                const float4 v = make_float4(s.x, s.y, s.x + k, s.y);
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    float4 *dst = reinterpret_cast<float4 *>(p.out[b] + off);
                    if (STMODE == 0) __stcs(dst, v);
                    else if (STMODE == 1) *dst = v;
                    else __stwt(dst, v);
                }
            }
        }
    }
}

static Prm g_prm;
static cudaEvent_t e0, e1;

template <typename K>
static void run(const char *name, K kern, int T1, int T2, int NT, size_t smem, double bytes, int cl = 1)
{
    Prm p = g_prm;
    p.tiles1 = (p.n1 + T1 - 1) / T1;
    p.tiles2 = (p.n2 + T2 - 1) / T2;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, smem));
    const unsigned grid = (unsigned)(p.tiles1 * p.tiles2 * p.nbatch);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        CK(cudaEventRecord(e0));
        if (cl > 1) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(grid);
            cfg.blockDim = dim3(NT);
            cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = (unsigned)cl;
            attr[0].val.clusterDim.y = 1;
            attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            CK(cudaLaunchKernelEx(&cfg, kern, p));
        } else
            kern<<<grid, NT, smem>>>(p);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    CK(cudaGetLastError());
    printf("%-44s T1=%3d T2=%2d NT=%3d smem=%6zu occ=%d  %7.3f ms  %7.1f GB/s (compulsory bytes)\n", name, T1, T2, NT, smem, occ,
           best, bytes / best * 1e-6);
}

template <int T1, int T2, int NT, int NBUF, int MINB, int CL = 1>
static void run_r(double bytes)
{
    constexpr int W1S = (T1 >= 192) ? T1 : T1 + 8, W2 = T2 + H;
    constexpr int BP = (W2 * W1S * 8 + 127) / 128 * 128 / 8;
    char name[64];
    snprintf(name, sizeof name, "probe_r nbuf=%d minb=%d cluster=%d", NBUF, MINB, CL);
    run(name, probe_r<T1, T2, NT, NBUF, MINB, CL>, T1, T2, NT, (size_t)NBUF * 8 * BP * 8 + 64, bytes, CL);
}
template <int T1, int T2, int NT, int MINB, int STMODE>
static void run_w(double bytes)
{
    char name[64];
    snprintf(name, sizeof name, "probe_w st=%s minb=%d", STMODE == 0 ? "cs" : STMODE == 1 ? "default" : "wt", MINB);
    run(name, probe_w<T1, T2, NT, MINB, STMODE>, T1, T2, NT, 0, bytes);
}

int main()
{
    const int n1 = 192, n2 = 192, n3 = 64, nbatch = 96;
    const int64_t s3 = (int64_t)n1 * n2, s4 = s3 * n3, N = s4 * nbatch;
    T *in, *out;
    CK(cudaMalloc(&in, (size_t)N * 8 * sizeof(T)));
    CK(cudaMalloc(&out, (size_t)N * 8 * sizeof(T)));
    CK(cudaMemset(in, 0, (size_t)N * 8 * sizeof(T)));
    CK(cudaMemset(out, 0, (size_t)N * 8 * sizeof(T)));
    for (int b = 0; b < 8; ++b) { g_prm.in[b] = in + (size_t)b * N; g_prm.out[b] = out + (size_t)b * N; }
    g_prm.n1 = n1; g_prm.n2 = n2; g_prm.n3 = n3; g_prm.nbatch = nbatch; g_prm.s3 = s3; g_prm.s4 = s4;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const double bytes = 9.0 * (double)N * 8.0;
    printf("cfg5 tile-kernel geometry: %dx%dx%d x %d batches, compulsory %.2f GB per launch\n", n1, n2, n3, nbatch, bytes * 1e-9);
    // read pattern (synthesis): T1, T2, NT, NBUF, MINB
    run_r<32, 16, 320, 1, 2>(bytes);     // what k_rec3_bulk does
    run_r<32, 16, 320, 1, 2, 6>(bytes);   // + lockstep cluster of the 6 dim-1 neighbours
    run_r<32, 16, 320, 1, 2, 3>(bytes);
    run_r<32, 16, 320, 1, 2, 2>(bytes);
    run_r<32, 8, 192, 1, 4, 6>(bytes);
    run_r<64, 8, 256, 1, 2, 3>(bytes);
    run_r<32, 16, 256, 2, 1>(bytes);
    run_r<32, 16, 256, 3, 1>(bytes);
    run_r<32, 8, 192, 1, 4>(bytes);
    run_r<32, 8, 192, 2, 2>(bytes);
    run_r<32, 8, 128, 3, 1>(bytes);
    run_r<64, 8, 256, 1, 2>(bytes);
    run_r<64, 8, 256, 2, 1>(bytes);
    run_r<64, 16, 512, 1, 1>(bytes);
    run_r<96, 8, 384, 1, 2>(bytes);
    run_r<96, 8, 384, 2, 1>(bytes);
    run_r<192, 8, 768, 1, 1>(bytes);
    run_r<192, 4, 384, 1, 2>(bytes);
    // write pattern (analysis): T1, T2, NT, MINB, store mode
    run_w<32, 16, 256, 2, 0>(bytes);     // what k_dec3_fused does
    run_w<32, 16, 256, 2, 1>(bytes);
    run_w<32, 16, 256, 2, 2>(bytes);
    run_w<32, 16, 256, 4, 0>(bytes);
    run_w<64, 8, 256, 2, 0>(bytes);
    run_w<64, 16, 512, 1, 0>(bytes);
    run_w<96, 8, 384, 2, 0>(bytes);
    run_w<192, 4, 384, 2, 0>(bytes);
    run_w<192, 8, 768, 1, 0>(bytes);
    run_w<192, 8, 768, 1, 1>(bytes);
    return 0;
}
