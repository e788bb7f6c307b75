// nddwt_multi.cu -- multi-GPU plan behind the C ABI (include/nddwt_b200.h, nddwt_mplan_*).
//
// The reference has no multi-device code; BASELINE.json's north_star asks for slabs along
// the LAST dimension with a halo exchange per level over NVLink.  Here the plan itself owns the peer
// mapping and the exchange, so that a C or MATLAB caller can drive all GPUs of a box:
//   * rank r owns a contiguous block of planes of the last dim of x and of every subband;
//   * per analysis level each rank PUSHES the (L/2-1) + L/2 approximation planes its neighbours need
//     straight into their halo inboxes with copy-engine peer copies (cudaMemcpyAsync on peer-mapped
//     memory: no SM is taken from the tile kernels, no staging through NCCL buffers), then raises a flag;
//   * synthesis is the adjoint: the last-dim pass runs in scatter form, the L-1 overhang planes of
//     partial sums are pushed to their owners, which add them (one array moves instead of 2^d bands);
//   * the tile pass that does not feed the next exchange (bands 2^(d-1)..2^d-1 in analysis, the
//     detail-only half of stage 1 in synthesis) runs while the planes are in flight.
// Two ways to use it, same schedule:
//   (1) nddwt_mplan_create: ONE process drives ngpus devices (peer access, CUDA events order the ranks);
//   (2) nddwt_mplan_create_rank + export/import: one process per GPU (torchrun); the halo inboxes are
//       shared through CUDA IPC handles and the ranks order themselves with flags in peer memory
//       (st.release.sys by the producer, an acquire spin with a time-out by the consumer).
// NCCL is not used on this path.
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "nddwt_plan.h"

namespace nddwt {

enum {
    CH_HALO_FILLED0 = 0, CH_HALO_FILLED1, CH_HALO_FREE0, CH_HALO_FREE1,   // analysis halo inboxes (two parities)
    CH_STAGE_FILLED, CH_STAGE_FREE,                                       // synthesis: overhang partial sums
    CH_RECH_FILLED, CH_RECH_FREE,                                         // synthesis, gather form: u halos
    NCH = 8
};

constexpr int EV_RING = 64;   // events mode: one event per signal sequence number (mod ring), so that a wait enqueued
                              // after several later signals still waits for ITS signal only

struct Run { int owner; int64_t idx, cnt, off; };   // planes idx.. of `owner` fill halo slots off..

struct Piece { int src, which; int64_t over_off, idx, cnt, stage_off; };   // overhang planes -> owner's planes idx..

struct RankCtx {
    int rank = 0, device = 0;
    nddwt_plan *plan = nullptr;
    int64_t start = 0, count = 0;
    cudaStream_t cs = nullptr;                 // plan-owned compute stream (used when the caller passes none)
    cudaStream_t ms = nullptr;                 // comm stream (high priority): flag waits / signals and copies
    cudaStream_t mx[3] = {nullptr, nullptr, nullptr};   // extra copy streams: every pushed run is cut in pieces, one per stream / copy engine
    cudaEvent_t ev_fork = nullptr, ev_join[3] = {nullptr, nullptr, nullptr};
    char *arena = nullptr;                     // peer-visible: flags, halo inboxes, stage
    void *approx[2] = {nullptr, nullptr};
    void *u_lo = nullptr, *u_hi[2] = {nullptr, nullptr};
    void *over[2] = {nullptr, nullptr};
    cudaEvent_t ev_prod = nullptr, ev_pushed = nullptr;
    bool pushed_once = false;
    void *host_x = nullptr, *host_c = nullptr;   // device staging of the *_host entry points (this rank's slabs)
    size_t host_c_bytes = 0;
    std::vector<uint32_t> sent, expect;        // [world * NCH] flag sequence numbers
    std::vector<uint32_t> rounds;              // [world * NCH] exchanges started towards a peer (FREE is raised once per exchange)
};

}  // namespace nddwt

using namespace nddwt;

struct nddwt_mplan {
    int world = 1, ndims = 0, dtype = 0, pres_l2 = 0;
    bool flags_mode = false;     // multi-process: IPC arenas + flags; otherwise CUDA events
    bool imported = false;
    int64_t dims[NDDWT_MAX_DIMS] = {1, 1, 1, 1};
    std::vector<std::string> wnames;
    int L_last = 2;
    size_t esize = 0;
    int64_t plane_elems = 1;
    size_t plane_bytes = 0;
    int dil[NDDWT_MAX_LEVELS];
    std::vector<int64_t> start, count;
    std::vector<RankCtx> local;
    std::vector<int> local_index;            // rank -> index in `local` or -1
    std::vector<char *> peer_arena;          // [world] arena base as seen from this process
    std::vector<bool> peer_opened;
    std::vector<cudaEvent_t> ev;             // events mode: [((a * world + b) * NCH + ch) * EV_RING + seq % EV_RING], recorded on a's stream
    std::vector<char> ev_recorded;
    bool separable = false;
    size_t off_flags = 0, off_inbox[2][2] = {{0, 0}, {0, 0}}, off_stage = 0, arena_bytes = 0;
    int64_t cap_lo = 0, cap_hi = 0, cap_stage = 0;   // inbox capacities (planes)
    uint32_t *status_host = nullptr;         // mapped host word: number of flag waits that timed out
    int64_t extra_launches = 0;
    int64_t copies = 0, copy_bytes = 0;
    int comm_streams = 1;                    // copy streams per rank (1..4), nddwt_mplan_set_param("comm_streams")
    int z_chunks = 0;                        // separable 4-D plans: a level is issued in this many dim-3 chunks so that
                                             // the halo planes of one chunk travel while the next chunk computes
                                             // (0 = chosen from the halo : slab ratio, see make_chunks)
};

namespace nddwt {

static void partition(int64_t n, int world, std::vector<int64_t> &start, std::vector<int64_t> &count)
{
    const int64_t base = n / world, rem = n % world;
    start.resize(world);
    count.resize(world);
    int64_t s = 0;
    for (int r = 0; r < world; ++r) {
        count[r] = base + (r < rem ? 1 : 0);
        start[r] = s;
        s += count[r];
    }
}

static int owner_of(int64_t plane, const std::vector<int64_t> &start, const std::vector<int64_t> &count, int64_t *local)
{
    for (size_t r = 0; r < start.size(); ++r)
        if (plane >= start[r] && plane < start[r] + count[r]) { *local = plane - start[r]; return (int)r; }
    *local = 0;
    return -1;
}

// planes a rank needs around its slab (which = 0: `below` planes under it, 1: `above` planes over it), periodic,
// resolved to (owner, first local plane, count, first halo slot) runs in ascending slot order
static std::vector<Run> halo_runs(int64_t n, const std::vector<int64_t> &start, const std::vector<int64_t> &count,
                                  int rank, int which, int64_t below, int64_t above)
{
    std::vector<Run> runs;
    const int64_t m = which ? above : below;
    for (int64_t i = 0; i < m; ++i) {
        int64_t g = which ? start[rank] + count[rank] + i : start[rank] - below + i;
        g %= n;
        if (g < 0) g += n;
        int64_t loc;
        const int o = owner_of(g, start, count, &loc);
        if (!runs.empty() && runs.back().owner == o && runs.back().idx + runs.back().cnt == loc) runs.back().cnt++;
        else runs.push_back({o, loc, 1, i});
    }
    return runs;
}

__global__ void k_flag_set(uint32_t *flag, uint32_t v)
{
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(v) : "memory");
}

// acquire spin on a flag in LOCAL memory that a peer GPU raises; gives up after `timeout_ns` and counts the
// failure in a host-mapped word instead of hanging the device
__global__ void k_flag_wait(const uint32_t *flag, uint32_t v, uint32_t *status, unsigned long long timeout_ns)
{
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        uint32_t cur;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(cur) : "l"(flag) : "memory");
        if ((int32_t)(cur - v) >= 0) return;
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (t - t0 > timeout_ns) { atomicAdd_system(status, 1u); return; }
        __nanosleep(200);
    }
}

static inline uint32_t *flag_ptr(const nddwt_mplan *mp, int at_rank, int from_rank, int ch)
{
    return reinterpret_cast<uint32_t *>(mp->peer_arena[at_rank] + mp->off_flags) + (size_t)from_rank * NCH + ch;
}

// rank `from` (local) tells rank `to` that channel `ch` advanced; ordered after everything queued on `s`
static int fab_signal(nddwt_mplan *mp, RankCtx &from, int to, int ch, cudaStream_t s)
{
    NDDWT_CUDA(cudaSetDevice(from.device));
    const uint32_t v = ++from.sent[(size_t)to * NCH + ch];
    if (mp->flags_mode) {
        k_flag_set<<<1, 1, 0, s>>>(flag_ptr(mp, to, from.rank, ch), v);
        mp->extra_launches++;
        NDDWT_CUDA(cudaGetLastError());
    } else {
        const size_t e = (((size_t)from.rank * mp->world + to) * NCH + ch) * EV_RING + v % EV_RING;
        if (!mp->ev[e]) NDDWT_CUDA(cudaEventCreateWithFlags(&mp->ev[e], cudaEventDisableTiming));
        NDDWT_CUDA(cudaEventRecord(mp->ev[e], s));
        mp->ev_recorded[e] = 1;
    }
    return 0;
}

// stream `s` of local rank `at` waits until rank `from` has signalled channel `ch` `value` times
// (value 0: the next signal in sequence)
static int fab_wait(nddwt_mplan *mp, RankCtx &at, int from, int ch, cudaStream_t s, uint32_t value = 0)
{
    NDDWT_CUDA(cudaSetDevice(at.device));
    uint32_t &exp = at.expect[(size_t)from * NCH + ch];
    if (value == 0) value = ++exp;
    else exp = value;
    if (mp->flags_mode) {
        k_flag_wait<<<1, 1, 0, s>>>(flag_ptr(mp, at.rank, from, ch), value, mp->status_host, 20ull * 1000 * 1000 * 1000);
        mp->extra_launches++;
        NDDWT_CUDA(cudaGetLastError());
    } else {
        const size_t e = (((size_t)from * mp->world + at.rank) * NCH + ch) * EV_RING + value % EV_RING;
        if (mp->ev_recorded[e]) NDDWT_CUDA(cudaStreamWaitEvent(s, mp->ev[e], 0));
    }
    return 0;
}

static inline char *inbox_ptr(const nddwt_mplan *mp, int rank, int par, int which)
{
    return mp->peer_arena[rank] + mp->off_inbox[par][which];
}

// one logical peer copy, cut into comm_streams pieces that run on different streams (copy engines) at once.
// The caller brackets a group of copies with fork_streams / join_streams on the rank's main comm stream.
static int peer_copy(nddwt_mplan *mp, RankCtx &r, char *dst, const char *src, size_t bytes)
{
    const int ns = mp->comm_streams;
    const size_t piece = ((bytes + ns - 1) / ns + 255) / 256 * 256;
    for (int i = 0; i < ns; ++i) {
        const size_t o = (size_t)i * piece;
        if (o >= bytes) break;
        const size_t n = bytes - o < piece ? bytes - o : piece;
        NDDWT_CUDA(cudaMemcpyAsync(dst + o, src + o, n, cudaMemcpyDefault, i == 0 ? r.ms : r.mx[i - 1]));
    }
    mp->copies++;
    mp->copy_bytes += (int64_t)bytes;
    return 0;
}
static int fork_streams(nddwt_mplan *mp, RankCtx &r)   // extra streams continue after what is queued on ms
{
    if (mp->comm_streams <= 1) return 0;
    NDDWT_CUDA(cudaEventRecord(r.ev_fork, r.ms));
    for (int i = 1; i < mp->comm_streams; ++i) NDDWT_CUDA(cudaStreamWaitEvent(r.mx[i - 1], r.ev_fork, 0));
    return 0;
}
static int join_streams(nddwt_mplan *mp, RankCtx &r)   // ms continues after the pieces on the extra streams
{
    for (int i = 1; i < mp->comm_streams; ++i) {
        NDDWT_CUDA(cudaEventRecord(r.ev_join[i - 1], r.mx[i - 1]));
        NDDWT_CUDA(cudaStreamWaitEvent(r.ms, r.ev_join[i - 1], 0));
    }
    return 0;
}

struct PushSrc { const char *base; int64_t slot_shift[2]; };   // local array + extra slot offset per side (gather-form u_hi)

// Rank r pushes, for every rank q, the planes of its local arrays that q's halo needs (halo shape below/above)
// into q's inbox `par`; one FREE wait before and one FILLED signal after per destination.
static int push_halos(nddwt_mplan *mp, RankCtx &r, const PushSrc *srcs, int nsrc, int64_t below, int64_t above, int par,
                      int ch_filled, int ch_free)
{
    NDDWT_CUDA(cudaSetDevice(r.device));
    const int64_t n = mp->dims[mp->ndims - 1];
    LaunchTimer lt(r.plan, KIND_COMM, r.ms);   // profiling: the whole push of this level on the comm stream
    for (int dq = 0; dq < mp->world; ++dq) {
        const int q = (r.rank + dq) % mp->world;
        bool any = false;
        for (int which = 0; which < 2; ++which) {
            const std::vector<Run> runs = halo_runs(n, mp->start, mp->count, q, which, below, above);
            for (const Run &run : runs) {
                if (run.owner != r.rank) continue;
                if (!any) {
                    if (q != r.rank) {
                        uint32_t &rounds = r.rounds[(size_t)q * NCH + ch_filled];
                        if (rounds > 0) { int rc = fab_wait(mp, r, q, ch_free, r.ms, rounds); if (rc) return rc; }
                        ++rounds;
                    }
                    NDDWT_CUDA(cudaSetDevice(r.device));
                    int rc = fork_streams(mp, r);
                    if (rc) return rc;
                }
                any = true;
                for (int a = 0; a < nsrc; ++a) {
                    char *dst = inbox_ptr(mp, q, par, which) + (size_t)(run.off + srcs[a].slot_shift[which]) * mp->plane_bytes;
                    const char *src = srcs[a].base + (size_t)run.idx * mp->plane_bytes;
                    int rc = peer_copy(mp, r, dst, src, (size_t)run.cnt * mp->plane_bytes);
                    if (rc) return rc;
                }
            }
        }
        if (any) { int rc = join_streams(mp, r); if (rc) return rc; }
        if (any && q != r.rank) { int rc = fab_signal(mp, r, q, ch_filled, r.ms); if (rc) return rc; }
    }
    NDDWT_CUDA(cudaSetDevice(r.device));
    NDDWT_CUDA(cudaEventRecord(r.ev_pushed, r.ms));
    r.pushed_once = true;
    return 0;
}

// ranks whose planes appear in the halo of `rank` (other than itself)
static std::vector<int> halo_sources_of(const nddwt_mplan *mp, int rank, int64_t below, int64_t above)
{
    std::vector<char> seen(mp->world, 0);
    const int64_t n = mp->dims[mp->ndims - 1];
    for (int which = 0; which < 2; ++which)
        for (const Run &run : halo_runs(n, mp->start, mp->count, rank, which, below, above)) seen[run.owner] = 1;
    std::vector<int> out;
    for (int q = 0; q < mp->world; ++q)
        if (seen[q] && q != rank) out.push_back(q);
    return out;
}

static int wait_halos(nddwt_mplan *mp, RankCtx &r, cudaStream_t cs, int64_t below, int64_t above, int ch_filled)
{
    for (int q : halo_sources_of(mp, r.rank, below, above)) { int rc = fab_wait(mp, r, q, ch_filled, cs); if (rc) return rc; }
    NDDWT_CUDA(cudaSetDevice(r.device));
    NDDWT_CUDA(cudaStreamWaitEvent(cs, r.ev_pushed, 0));   // own (wrap-around) planes are copied locally on ms
    return 0;
}

static int free_halos(nddwt_mplan *mp, RankCtx &r, cudaStream_t cs, int64_t below, int64_t above, int ch_free)
{
    for (int q : halo_sources_of(mp, r.rank, below, above)) { int rc = fab_signal(mp, r, q, ch_free, cs); if (rc) return rc; }
    return 0;
}

// scatter-form synthesis: the overhang partial sums every rank `src` computes for planes of `owner`, in one
// deterministic order that fixes the layout of the owner's stage buffer (both sides derive it)
static std::vector<Piece> pieces_into(const nddwt_mplan *mp, int owner, int64_t below, int64_t above)
{
    std::vector<Piece> out;
    const int64_t n = mp->dims[mp->ndims - 1];
    int64_t off = 0;
    for (int src = 0; src < mp->world; ++src)
        for (int which = 0; which < 2; ++which)
            for (const Run &run : halo_runs(n, mp->start, mp->count, src, which, below, above)) {
                if (run.owner != owner) continue;
                Piece p;
                p.src = src; p.which = which; p.over_off = run.off; p.idx = run.idx; p.cnt = run.cnt;
                p.stage_off = (src == owner) ? -1 : off;
                if (src != owner) off += run.cnt;
                out.push_back(p);
            }
    return out;
}

static int64_t stage_planes_of(const nddwt_mplan *mp, int owner, int64_t below, int64_t above)
{
    int64_t tot = 0;
    for (const Piece &p : pieces_into(mp, owner, below, above))
        if (p.src != owner) tot += p.cnt;
    return tot;
}

static void release(nddwt_mplan *mp)
{
    if (!mp) return;
    for (RankCtx &c : mp->local) {
        cudaSetDevice(c.device);
        if (c.plan) nddwt_plan_destroy(c.plan);
        for (int i = 0; i < 2; ++i) { if (c.approx[i]) cudaFree(c.approx[i]); if (c.u_hi[i]) cudaFree(c.u_hi[i]); if (c.over[i]) cudaFree(c.over[i]); }
        if (c.u_lo) cudaFree(c.u_lo);
        if (c.host_x) cudaFree(c.host_x);
        if (c.host_c) cudaFree(c.host_c);
        if (c.arena) cudaFree(c.arena);
        if (c.cs) cudaStreamDestroy(c.cs);
        if (c.ms) cudaStreamDestroy(c.ms);
        for (int i = 0; i < 3; ++i) { if (c.mx[i]) cudaStreamDestroy(c.mx[i]); if (c.ev_join[i]) cudaEventDestroy(c.ev_join[i]); }
        if (c.ev_fork) cudaEventDestroy(c.ev_fork);
        if (c.ev_prod) cudaEventDestroy(c.ev_prod);
        if (c.ev_pushed) cudaEventDestroy(c.ev_pushed);
    }
    if (mp->flags_mode && !mp->local.empty()) {
        cudaSetDevice(mp->local[0].device);
        for (int q = 0; q < mp->world; ++q)
            if (mp->peer_opened[q] && mp->peer_arena[q]) cudaIpcCloseMemHandle(mp->peer_arena[q]);
    }
    for (cudaEvent_t e : mp->ev) if (e) cudaEventDestroy(e);
    if (mp->status_host) cudaFreeHost(mp->status_host);
    delete mp;
}

static int dev_alloc(void **p, size_t bytes)
{
    NDDWT_CUDA(cudaMalloc(p, bytes ? bytes : 16));
    return 0;
}

static int create_common(nddwt_mplan **out, int ndims, const int64_t *dims, const char *const *wnames, int dtype,
                         int pres_l2, int world, const std::vector<int> &ranks, const std::vector<int> &devices,
                         bool flags_mode)
{
    if (!out || !dims || !wnames) { set_error("null argument"); return NDDWT_ERR_ARG; }
    *out = nullptr;
    if (ndims < 1 || ndims > NDDWT_MAX_DIMS) { set_error("ndims must be 1..4"); return NDDWT_ERR_ARG; }
    if (world < 1 || dims[ndims - 1] < world) { set_error("the last dimension must have at least one plane per rank"); return NDDWT_ERR_ARG; }
    nddwt_mplan *mp = new nddwt_mplan();
    mp->world = world;
    mp->ndims = ndims;
    mp->dtype = dtype;
    mp->pres_l2 = pres_l2 ? 1 : 0;
    mp->flags_mode = flags_mode;
    for (int j = 0; j < NDDWT_MAX_LEVELS; ++j) mp->dil[j] = 1;
    for (int i = 0; i < ndims; ++i) { mp->dims[i] = dims[i]; mp->wnames.push_back(wnames[i] ? wnames[i] : ""); }
    partition(dims[ndims - 1], world, mp->start, mp->count);
    mp->plane_elems = 1;
    for (int i = 0; i + 1 < ndims; ++i) mp->plane_elems *= dims[i];
    mp->local_index.assign(world, -1);
    mp->peer_arena.assign(world, nullptr);
    mp->peer_opened.assign(world, false);
    mp->local.resize(ranks.size());
    int rc = 0;
    for (size_t i = 0; i < ranks.size() && !rc; ++i) {
        RankCtx &c = mp->local[i];
        c.rank = ranks[i];
        c.device = devices[i];
        c.start = mp->start[c.rank];
        c.count = mp->count[c.rank];
        c.sent.assign((size_t)world * NCH, 0);
        c.expect.assign((size_t)world * NCH, 0);
        c.rounds.assign((size_t)world * NCH, 0);
        mp->local_index[c.rank] = (int)i;
        int64_t ld[NDDWT_MAX_DIMS];
        for (int k = 0; k < ndims; ++k) ld[k] = dims[k];
        ld[ndims - 1] = c.count;
        rc = nddwt_plan_create_slab(&c.plan, ndims, ld, dims[ndims - 1], wnames, dtype, pres_l2, c.device);
    }
    if (rc) { release(mp); return rc; }
    mp->esize = mp->local[0].plan->esize;
    mp->plane_bytes = (size_t)mp->plane_elems * mp->esize;
    mp->L_last = mp->local[0].plan->L[ndims - 1];
    mp->separable = nddwt_plan_is_separable(mp->local[0].plan) != 0;
    for (RankCtx &c : mp->local) mp->separable = mp->separable && nddwt_plan_is_separable(c.plan) != 0;
    *out = mp;
    return 0;
}

// (re)allocates the per-rank buffers for the current dilations; arena layout is identical on every rank
static int allocate(nddwt_mplan *mp)
{
    int maxdil = 1;
    for (int j = 0; j < NDDWT_MAX_LEVELS; ++j) maxdil = mp->dil[j] > maxdil ? mp->dil[j] : maxdil;
    const int L = mp->L_last;
    // inboxes hold the analysis halos (L/2-1 | L/2 planes) and, in the gather-form synthesis, the u_lo and
    // u_hi halos (2 x L/2 | 2 x (L/2-1) planes); one capacity covers both
    mp->cap_lo = (int64_t)2 * (L / 2) * maxdil;
    mp->cap_hi = (int64_t)2 * (L / 2) * maxdil;
    mp->cap_stage = 0;
    for (int q = 0; q < mp->world; ++q) {
        const int64_t s = stage_planes_of(mp, q, (int64_t)(L / 2 - 1) * maxdil, (int64_t)(L / 2) * maxdil);
        mp->cap_stage = s > mp->cap_stage ? s : mp->cap_stage;
    }
    size_t off = 0;
    mp->off_flags = off;
    off += ((size_t)mp->world * NCH * sizeof(uint32_t) + 255) / 256 * 256;
    for (int par = 0; par < 2; ++par) {
        mp->off_inbox[par][0] = off; off += (size_t)mp->cap_lo * mp->plane_bytes; off = (off + 255) / 256 * 256;
        mp->off_inbox[par][1] = off; off += (size_t)mp->cap_hi * mp->plane_bytes; off = (off + 255) / 256 * 256;
    }
    mp->off_stage = off;
    off += (size_t)mp->cap_stage * mp->plane_bytes;
    mp->arena_bytes = (off + 255) / 256 * 256;
    for (RankCtx &c : mp->local) {
        NDDWT_CUDA(cudaSetDevice(c.device));
        const size_t slab_bytes = (size_t)c.count * mp->plane_bytes;
        int rc = dev_alloc(reinterpret_cast<void **>(&c.arena), mp->arena_bytes);
        if (rc) return rc;
        NDDWT_CUDA(cudaMemset(c.arena, 0, mp->off_inbox[0][0]));
        for (int i = 0; i < 2 && !rc; ++i) rc = dev_alloc(&c.approx[i], slab_bytes);
        if (!rc) rc = dev_alloc(&c.u_lo, slab_bytes);
        for (int i = 0; i < 2 && !rc; ++i) rc = dev_alloc(&c.u_hi[i], slab_bytes);
        if (!rc) rc = dev_alloc(&c.over[0], (size_t)(L / 2 - 1) * maxdil * mp->plane_bytes);
        if (!rc) rc = dev_alloc(&c.over[1], (size_t)(L / 2) * maxdil * mp->plane_bytes);
        if (rc) return rc;
        int lo_prio = 0, hi_prio = 0;
        NDDWT_CUDA(cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio));
        NDDWT_CUDA(cudaStreamCreateWithFlags(&c.cs, cudaStreamNonBlocking));
        NDDWT_CUDA(cudaStreamCreateWithPriority(&c.ms, cudaStreamNonBlocking, hi_prio));
        for (int i = 0; i < 3; ++i) {
            NDDWT_CUDA(cudaStreamCreateWithPriority(&c.mx[i], cudaStreamNonBlocking, hi_prio));
            NDDWT_CUDA(cudaEventCreateWithFlags(&c.ev_join[i], cudaEventDisableTiming));
        }
        NDDWT_CUDA(cudaEventCreateWithFlags(&c.ev_fork, cudaEventDisableTiming));
        NDDWT_CUDA(cudaEventCreateWithFlags(&c.ev_prod, cudaEventDisableTiming));
        NDDWT_CUDA(cudaEventCreateWithFlags(&c.ev_pushed, cudaEventDisableTiming));
        NDDWT_CUDA(cudaDeviceSynchronize());
        mp->peer_arena[c.rank] = c.arena;
    }
    NDDWT_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&mp->status_host), 64, cudaHostAllocMapped | cudaHostAllocPortable));
    memset(mp->status_host, 0, 64);
    if (!mp->flags_mode) {
        mp->ev.assign((size_t)mp->world * mp->world * NCH * EV_RING, nullptr);
        mp->ev_recorded.assign((size_t)mp->world * mp->world * NCH * EV_RING, 0);
    }
    return 0;
}

// ---- z-chunked exchange (separable 4-D plans) ---------------------------------------------------------
struct Chunks {
    int C = 1;
    std::vector<int> zb;          // chunk c covers dim-3 planes zb[c] .. zb[c+1]-1
    size_t row_bytes = 0;         // bytes of one dim-3 plane of one hyperplane (n1 * n2 elements)
    size_t off(int c) const { return (size_t)zb[c] * row_bytes; }
    size_t bytes(int c) const { return (size_t)(zb[c + 1] - zb[c]) * row_bytes; }
    ZRange range(int c) const { ZRange z; z.z0 = zb[c]; z.zn = zb[c + 1] - zb[c]; return z; }
};

static Chunks make_chunks(const nddwt_mplan *mp)
{
    Chunks ch;
    const int n3 = (int)mp->dims[2];
    int C = mp->z_chunks;
    if (C <= 0) {
        // Chunking hides the flight of the halo planes behind compute but costs extra kernel tails and ring warm-ups;
        // it pays when the halo is large against the slab (measured on cfg4, profiles/r02_multi_gpu.md: 4 planes per
        // rank 23.2 -> 19.5 ms with 4 chunks; 16 planes per rank 66.2 / 65.7 / 67.3 ms with 1 / 2 / 4 chunks)
        int64_t maxc = 1;
        for (int64_t c : mp->count) maxc = c > maxc ? c : maxc;
        const double ratio = (double)(mp->L_last - 1) / (double)maxc;
        C = ratio > 1.2 ? 4 : (ratio > 0.3 ? 2 : 1);
    }
    if (C > n3 / 8) C = n3 / 8;      // a chunk must be wider than the dim-3 filter support
    if (C < 1) C = 1;
    ch.C = C;
    for (int c = 0; c <= C; ++c) ch.zb.push_back((int)((int64_t)c * n3 / C));
    ch.row_bytes = (size_t)mp->dims[0] * mp->dims[1] * mp->esize;
    return ch;
}

// `height` hyperplanes, `width` bytes of each (one z-chunk), cut in comm_streams pieces
static int peer_copy2d(nddwt_mplan *mp, RankCtx &r, char *dst, const char *src, size_t width, size_t height, cudaStream_t only)
{
    const int ns = only ? 1 : mp->comm_streams;
    const size_t piece = ((width + ns - 1) / ns + 255) / 256 * 256;
    for (int i = 0; i < ns; ++i) {
        const size_t o = (size_t)i * piece;
        if (o >= width) break;
        const size_t n = width - o < piece ? width - o : piece;
        cudaStream_t st = only ? only : (i == 0 ? r.ms : r.mx[i - 1]);
        NDDWT_CUDA(cudaMemcpy2DAsync(dst + o, mp->plane_bytes, src + o, mp->plane_bytes, n, height, cudaMemcpyDefault, st));
    }
    if (!only) { mp->copies++; mp->copy_bytes += (int64_t)(width * height); }
    return 0;
}

// rank r pushes chunk c of the planes of `base` its neighbours need (analysis halo shape) into their inboxes `par`
static int push_halo_chunk(nddwt_mplan *mp, RankCtx &r, const char *base, const Chunks &ch, int c, bool first, int par)
{
    NDDWT_CUDA(cudaSetDevice(r.device));
    const int64_t n = mp->dims[mp->ndims - 1];
    const int L = mp->L_last;
    const int64_t below = L / 2 - 1, above = L / 2;
    const int ch_filled = CH_HALO_FILLED0 + par, ch_free = CH_HALO_FREE0 + par;
    LaunchTimer lt(r.plan, KIND_COMM, r.ms);
    for (int dq = 1; dq < mp->world; ++dq) {
        const int q = (r.rank + dq) % mp->world;
        bool any = false;
        for (int which = 0; which < 2; ++which)
            for (const Run &run : halo_runs(n, mp->start, mp->count, q, which, below, above)) {
                if (run.owner != r.rank) continue;
                if (!any) {
                    if (first) {
                        uint32_t &rounds = r.rounds[(size_t)q * NCH + ch_filled];
                        if (rounds > 0) { int rc = fab_wait(mp, r, q, ch_free, r.ms, rounds); if (rc) return rc; }
                        ++rounds;
                    }
                    NDDWT_CUDA(cudaSetDevice(r.device));
                    int rc = fork_streams(mp, r);
                    if (rc) return rc;
                    any = true;
                }
                int rc = peer_copy2d(mp, r, inbox_ptr(mp, q, par, which) + (size_t)run.off * mp->plane_bytes + ch.off(c),
                                     base + (size_t)run.idx * mp->plane_bytes + ch.off(c), ch.bytes(c), (size_t)run.cnt, nullptr);
                if (rc) return rc;
            }
        if (any) {
            int rc = join_streams(mp, r);
            if (rc) return rc;
            rc = fab_signal(mp, r, q, ch_filled, r.ms);
            if (rc) return rc;
        }
    }
    return 0;
}

// planes of the rank's own slab that wrap around into its own halo (few ranks, thin slabs): local copies on the compute stream
static int self_halo_chunk(nddwt_mplan *mp, RankCtx &r, const char *base, const Chunks &ch, int c, int par, cudaStream_t cs)
{
    const int64_t n = mp->dims[mp->ndims - 1];
    const int L = mp->L_last;
    for (int which = 0; which < 2; ++which)
        for (const Run &run : halo_runs(n, mp->start, mp->count, r.rank, which, L / 2 - 1, L / 2)) {
            if (run.owner != r.rank) continue;
            NDDWT_CUDA(cudaSetDevice(r.device));
            int rc = peer_copy2d(mp, r, inbox_ptr(mp, r.rank, par, which) + (size_t)run.off * mp->plane_bytes + ch.off(c),
                                 base + (size_t)run.idx * mp->plane_bytes + ch.off(c), ch.bytes(c), (size_t)run.cnt, cs);
            if (rc) return rc;
        }
    return 0;
}

static void band_ptrs(const nddwt_mplan *mp, const RankCtx &c, char *coeffs, int level, int j, void **bands);

static int dec_chunked(nddwt_mplan *mp, const void *const *x_slabs, void *const *coeff_slabs, int level,
                       const std::vector<cudaStream_t> &cs, const Chunks &ch)
{
    const int nl = (int)mp->local.size(), nd = 1 << mp->ndims, L = mp->L_last, C = ch.C;
    const int64_t below = L / 2 - 1, above = L / 2;
    int rc = 0;
    std::vector<const void *> a_in(x_slabs, x_slabs + nl);
    std::vector<std::vector<void *>> bands(nl, std::vector<void *>(nd));
    for (int j = 1; j <= level; ++j) {
        const int par = j & 1;
        std::vector<int> order(C);
        for (int k = 0; k < C; ++k) order[k] = ((j - 1) + k) % C;   // chunks arrive in the order the previous level produced them
        for (int i = 0; i < nl; ++i) {
            RankCtx &c = mp->local[i];
                NDDWT_CUDA(cudaSetDevice(c.device));
            band_ptrs(mp, c, reinterpret_cast<char *>(coeff_slabs[i]), level, j, bands[i].data());
            bands[i][0] = (j == level) ? coeff_slabs[i] : c.approx[j & 1];
            c.plan->cur_level = j;
            if (j == 1) {      // x is complete: all its chunks can leave at once
                NDDWT_CUDA(cudaSetDevice(c.device));
                NDDWT_CUDA(cudaEventRecord(c.ev_prod, cs[i]));
                NDDWT_CUDA(cudaStreamWaitEvent(c.ms, c.ev_prod, 0));
                for (int k = 0; k < C; ++k) {
                    rc = push_halo_chunk(mp, c, reinterpret_cast<const char *>(a_in[i]), ch, order[k], k == 0, par);
                    if (rc) return rc;
                }
            }
        }
        auto P1 = [&](int k) -> int {      // last-dim pass of chunk order[k]: the only reader of the halos
            for (int i = 0; i < nl; ++i) {
                RankCtx &c = mp->local[i];
                NDDWT_CUDA(cudaSetDevice(c.device));
                for (int q : halo_sources_of(mp, c.rank, below, above)) { int r2 = fab_wait(mp, c, q, CH_HALO_FILLED0 + par, cs[i]); if (r2) return r2; }
                int r2 = self_halo_chunk(mp, c, reinterpret_cast<const char *>(a_in[i]), ch, order[k], par, cs[i]);
                if (r2) return r2;
                LevelIO io;
                io.halo_lo = inbox_ptr(mp, c.rank, par, 0);
                io.halo_hi = inbox_ptr(mp, c.rank, par, 1);
                r2 = fused_dec_level_part(c.plan, 1, 1, a_in[i], io, bands[i].data(), cs[i], ch.range(order[k]));
                if (r2) return r2 > 0 ? NDDWT_ERR_ARG : r2;
                if (k == C - 1) { r2 = free_halos(mp, c, cs[i], below, above, CH_HALO_FREE0 + par); if (r2) return r2; }
            }
            return 0;
        };
        auto P2 = [&](int k, bool first_push) -> int {   // tile pass of the lo half (incl. a_j) for chunk order[k]; a_j's chunk leaves at once
            for (int i = 0; i < nl; ++i) {
                RankCtx &c = mp->local[i];
                NDDWT_CUDA(cudaSetDevice(c.device));
                LevelIO none;
                int r2 = fused_dec_level_part(c.plan, 1, 2, a_in[i], none, bands[i].data(), cs[i], ch.range(order[k]));
                if (r2) return r2 > 0 ? NDDWT_ERR_ARG : r2;
                if (j < level) {
                    NDDWT_CUDA(cudaSetDevice(c.device));
                    NDDWT_CUDA(cudaEventRecord(c.ev_prod, cs[i]));
                    NDDWT_CUDA(cudaStreamWaitEvent(c.ms, c.ev_prod, 0));
                    r2 = push_halo_chunk(mp, c, reinterpret_cast<const char *>(bands[i][0]), ch, order[k], first_push, par ^ 1);
                    if (r2) return r2;
                }
            }
            return 0;
        };
        // software pipeline: the tile pass of a chunk needs the last-dim pass of both neighbouring chunks
        if ((rc = P1(0))) return rc;
        if (C > 1 && (rc = P1(1))) return rc;
        for (int k = 1; k < C; ++k) {
            if (k + 1 < C && (rc = P1(k + 1))) return rc;
            if ((rc = P2(k, k == 1))) return rc;
        }
        if ((rc = P2(0, C == 1))) return rc;
        for (int i = 0; i < nl; ++i) {      // remaining detail bands (hi half), whole range, while a_j travels
            RankCtx &c = mp->local[i];
                NDDWT_CUDA(cudaSetDevice(c.device));
            LevelIO none;
            rc = fused_dec_level_part(c.plan, 1, 3, a_in[i], none, bands[i].data(), cs[i]);
            if (rc) return rc > 0 ? NDDWT_ERR_ARG : rc;
            c.plan->last_path = 1;
            a_in[i] = bands[i][0];
        }
    }
    return 0;
}

static int rec_chunked(nddwt_mplan *mp, const void *const *coeff_slabs, void *const *x_slabs, int level,
                       const std::vector<cudaStream_t> &cs, const Chunks &ch)
{
    const int nl = (int)mp->local.size(), nd = 1 << mp->ndims, L = mp->L_last, C = ch.C;
    const int64_t below = L / 2 - 1, above = L / 2;
    int rc = 0;
    std::vector<const void *> a(coeff_slabs, coeff_slabs + nl);
    std::vector<std::vector<void *>> bands(nl, std::vector<void *>(nd));
    auto set_bands = [&](int i, int j, const void *approx) {
        band_ptrs(mp, mp->local[i], reinterpret_cast<char *>(const_cast<void *>(coeff_slabs[i])), level, j, bands[i].data());
        bands[i][0] = const_cast<void *>(approx);
    };
    for (int i = 0; i < nl; ++i) {
        RankCtx &c = mp->local[i];
                NDDWT_CUDA(cudaSetDevice(c.device));
        set_bands(i, level, a[i]);
        rc = fused_rec_stage1(c.plan, 1, bands[i].data(), c.u_lo, c.u_hi[level & 1], cs[i], 2);
        if (rc) return rc > 0 ? NDDWT_ERR_ARG : rc;
    }
    for (int j = level; j >= 1; --j) {
        const int k2 = j & 1;
        std::vector<void *> dst(nl);
        std::vector<std::vector<Piece>> pieces(nl);
        for (int i = 0; i < nl; ++i) {
            RankCtx &c = mp->local[i];
                NDDWT_CUDA(cudaSetDevice(c.device));
            dst[i] = (j == 1) ? x_slabs[i] : c.approx[j & 1];
            set_bands(i, j, a[i]);
            pieces[i] = pieces_into(mp, c.rank, below, above);
            NDDWT_CUDA(cudaSetDevice(c.device));
            if (c.pushed_once) NDDWT_CUDA(cudaStreamWaitEvent(cs[i], c.ev_pushed, 0));   // the previous level's overhangs have left
        }
        auto adds = [&](int kc) -> int {      // add what the neighbours computed for chunk kc of my planes
            for (int i = 0; i < nl; ++i) {
                RankCtx &c = mp->local[i];
                NDDWT_CUDA(cudaSetDevice(c.device));
                std::vector<char> from(mp->world, 0);
                for (const Piece &p : pieces[i]) if (p.src != c.rank) from[p.src] = 1;
                for (int q = 0; q < mp->world; ++q)
                    if (from[q]) { int r2 = fab_wait(mp, c, q, CH_STAGE_FILLED, cs[i]); if (r2) return r2; }
                AccParams acc;
                acc.n = 0;
                std::vector<int> slot((size_t)c.count, -1);
                bool fits = true;
                for (const Piece &p : pieces[i]) {
                    const char *s = (p.src == c.rank)
                                        ? reinterpret_cast<const char *>(c.over[p.which]) + (size_t)p.over_off * mp->plane_bytes
                                        : c.arena + mp->off_stage + (size_t)p.stage_off * mp->plane_bytes;
                    for (int64_t t = 0; t < p.cnt && fits; ++t) {
                        int &k = slot[p.idx + t];
                        if (k < 0) {
                            if (acc.n >= ACC_MAXP) { fits = false; break; }
                            k = acc.n++;
                            acc.item[k].dst = reinterpret_cast<char *>(dst[i]) + (size_t)(p.idx + t) * mp->plane_bytes + ch.off(kc);
                            acc.item[k].ns = 0;
                        }
                        if (acc.item[k].ns >= ACC_MAXS) { fits = false; break; }
                        acc.item[k].src[acc.item[k].ns++] = s + (size_t)t * mp->plane_bytes + ch.off(kc);
                    }
                }
                NDDWT_CUDA(cudaSetDevice(c.device));
                int r2 = 0;
                if (fits) {
                    r2 = accumulate_planes(c.plan, acc, (int64_t)(ch.bytes(kc) / mp->esize), cs[i]);
                    if (r2) return r2;
                } else {      // very long filters on very thin slabs: one add per overhang plane
                    for (const Piece &p : pieces[i]) {
                        const char *s = (p.src == c.rank)
                                            ? reinterpret_cast<const char *>(c.over[p.which]) + (size_t)p.over_off * mp->plane_bytes
                                            : c.arena + mp->off_stage + (size_t)p.stage_off * mp->plane_bytes;
                        for (int64_t t = 0; t < p.cnt; ++t) {
                            r2 = nddwt_accumulate(c.plan, reinterpret_cast<char *>(dst[i]) + (size_t)(p.idx + t) * mp->plane_bytes + ch.off(kc),
                                                  s + (size_t)t * mp->plane_bytes + ch.off(kc), (int64_t)(ch.bytes(kc) / mp->esize), cs[i]);
                            if (r2) return r2;
                        }
                    }
                }
                if (kc == C - 1)
                    for (int q = 0; q < mp->world; ++q)
                        if (from[q]) { r2 = fab_signal(mp, c, q, CH_STAGE_FREE, cs[i]); if (r2) return r2; }
            }
            return 0;
        };
        for (int k = 0; k < C; ++k) {
            for (int i = 0; i < nl; ++i) {
                RankCtx &c = mp->local[i];
                NDDWT_CUDA(cudaSetDevice(c.device));
                rc = fused_rec_stage1(c.plan, 1, bands[i].data(), c.u_lo, c.u_hi[k2], cs[i], 1, ch.range(k));
                if (rc) return rc > 0 ? NDDWT_ERR_ARG : rc;
                rc = fused_rec_stage2_scatter(c.plan, 1, c.u_lo, c.u_hi[k2], dst[i], c.over[0], c.over[1], cs[i], ch.range(k));
                if (rc) return rc > 0 ? NDDWT_ERR_ARG : rc;
                NDDWT_CUDA(cudaSetDevice(c.device));
                NDDWT_CUDA(cudaEventRecord(c.ev_prod, cs[i]));
                NDDWT_CUDA(cudaStreamWaitEvent(c.ms, c.ev_prod, 0));
                LaunchTimer lt(c.plan, KIND_COMM, c.ms);
                for (int dq = 1; dq < mp->world; ++dq) {
                    const int q = (c.rank + dq) % mp->world;
                    bool any = false;
                    for (const Piece &p : pieces_into(mp, q, below, above)) {
                        if (p.src != c.rank) continue;
                        if (!any) {
                            if (k == 0) {
                                uint32_t &rounds = c.rounds[(size_t)q * NCH + CH_STAGE_FILLED];
                                if (rounds > 0) { rc = fab_wait(mp, c, q, CH_STAGE_FREE, c.ms, rounds); if (rc) return rc; }
                                ++rounds;
                            }
                            NDDWT_CUDA(cudaSetDevice(c.device));
                            rc = fork_streams(mp, c);
                            if (rc) return rc;
                            any = true;
                        }
                        rc = peer_copy2d(mp, c, mp->peer_arena[q] + mp->off_stage + (size_t)p.stage_off * mp->plane_bytes + ch.off(k),
                                         reinterpret_cast<const char *>(c.over[p.which]) + (size_t)p.over_off * mp->plane_bytes + ch.off(k),
                                         ch.bytes(k), (size_t)p.cnt, nullptr);
                        if (rc) return rc;
                    }
                    if (any) {
                        rc = join_streams(mp, c);
                        if (rc) return rc;
                        rc = fab_signal(mp, c, q, CH_STAGE_FILLED, c.ms);
                        if (rc) return rc;
                    }
                }
                if (k == C - 1) {
                    NDDWT_CUDA(cudaSetDevice(c.device));
                    NDDWT_CUDA(cudaEventRecord(c.ev_pushed, c.ms));
                    c.pushed_once = true;
                }
            }
            if (k >= 1 && (rc = adds(k - 1))) return rc;
        }
        if (j > 1)      // detail-only half of the next level: independent of this level's result, hides the last chunk's flight
            for (int i = 0; i < nl; ++i) {
                RankCtx &c = mp->local[i];
                NDDWT_CUDA(cudaSetDevice(c.device));
                set_bands(i, j - 1, coeff_slabs[i]);
                rc = fused_rec_stage1(c.plan, 1, bands[i].data(), c.u_lo, c.u_hi[(j - 1) & 1], cs[i], 2);
                if (rc) return rc > 0 ? NDDWT_ERR_ARG : rc;
            }
        if ((rc = adds(C - 1))) return rc;
        for (int i = 0; i < nl; ++i) { a[i] = dst[i]; mp->local[i].plan->last_path = 1; }
    }
    return 0;
}

static void band_ptrs(const nddwt_mplan *mp, const RankCtx &c, char *coeffs, int level, int j, void **bands)
{
    const int nd = 1 << mp->ndims;
    const size_t band_bytes = (size_t)c.count * mp->plane_bytes;
    const int64_t start = (int64_t)(nd - 1) * (level - j);   // slot arithmetic of mex/nddwt.c:209-210,226
    for (int b = 1; b < nd; ++b) bands[b] = coeffs + (size_t)(start + b) * band_bytes;
}

struct ExportBlob {
    uint32_t magic;
    int32_t rank, device, pad;
    uint64_t arena_bytes;
    cudaIpcMemHandle_t handle;
};

}  // namespace nddwt

extern "C" {

int nddwt_mplan_create(nddwt_mplan **mp_out, int ndims, const int64_t *dims, const char *const *wnames, int dtype,
                       int pres_l2_norm, int ngpus, const int *devices)
{
    if (ngpus < 1 || ngpus > 64) { set_error("ngpus must be 1..64"); return NDDWT_ERR_ARG; }
    std::vector<int> ranks(ngpus), devs(ngpus);
    for (int i = 0; i < ngpus; ++i) { ranks[i] = i; devs[i] = devices ? devices[i] : i; }
    nddwt_mplan *mp = nullptr;
    int rc = create_common(&mp, ndims, dims, wnames, dtype, pres_l2_norm, ngpus, ranks, devs, false);
    if (rc) return rc;
    rc = allocate(mp);
    if (rc) { release(mp); return rc; }
    // direct peer copies between distinct devices (ignored where the topology does not allow it:
    // cudaMemcpyAsync then stages through the host)
    for (int a = 0; a < ngpus; ++a)
        for (int b = 0; b < ngpus; ++b) {
            if (devs[a] == devs[b]) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, devs[a], devs[b]) == cudaSuccess && can) {
                cudaSetDevice(devs[a]);
                cudaError_t e = cudaDeviceEnablePeerAccess(devs[b], 0);
                if (e != cudaSuccess) cudaGetLastError();   // already enabled: fine
            }
        }
    mp->imported = true;
    *mp_out = mp;
    return 0;
}

int nddwt_mplan_create_rank(nddwt_mplan **mp_out, int ndims, const int64_t *dims, const char *const *wnames, int dtype,
                            int pres_l2_norm, int rank, int world, int device)
{
    if (world < 1 || rank < 0 || rank >= world) { set_error("bad rank / world"); return NDDWT_ERR_ARG; }
    nddwt_mplan *mp = nullptr;
    int rc = create_common(&mp, ndims, dims, wnames, dtype, pres_l2_norm, world, std::vector<int>(1, rank),
                           std::vector<int>(1, device), world > 1);
    if (rc) return rc;
    rc = allocate(mp);
    if (rc) { release(mp); return rc; }
    mp->imported = (world == 1);
    *mp_out = mp;
    return 0;
}

int64_t nddwt_mplan_export_size(void) { return (int64_t)sizeof(ExportBlob); }

int nddwt_mplan_export(nddwt_mplan *mp, void *blob)
{
    if (!mp || !blob || mp->local.size() != 1) { set_error("export needs a rank plan"); return NDDWT_ERR_ARG; }
    RankCtx &c = mp->local[0];
    ExportBlob b;
    memset(&b, 0, sizeof b);
    b.magic = 0x4e445754u;
    b.rank = c.rank;
    b.device = c.device;
    b.arena_bytes = mp->arena_bytes;
    NDDWT_CUDA(cudaSetDevice(c.device));
    if (mp->world > 1) NDDWT_CUDA(cudaIpcGetMemHandle(&b.handle, c.arena));
    memcpy(blob, &b, sizeof b);
    return 0;
}

int nddwt_mplan_import(nddwt_mplan *mp, const void *blobs)
{
    if (!mp || !blobs || mp->local.size() != 1) { set_error("import needs a rank plan"); return NDDWT_ERR_ARG; }
    RankCtx &c = mp->local[0];
    NDDWT_CUDA(cudaSetDevice(c.device));
    const ExportBlob *b = reinterpret_cast<const ExportBlob *>(blobs);
    for (int q = 0; q < mp->world; ++q) {
        if (b[q].magic != 0x4e445754u || b[q].rank != q || b[q].arena_bytes != mp->arena_bytes) {
            set_error("peer blob does not match this plan (rank order, geometry)");
            return NDDWT_ERR_ARG;
        }
        if (q == c.rank || mp->peer_opened[q]) continue;
        void *ptr = nullptr;
        NDDWT_CUDA(cudaIpcOpenMemHandle(&ptr, b[q].handle, cudaIpcMemLazyEnablePeerAccess));
        mp->peer_arena[q] = reinterpret_cast<char *>(ptr);
        mp->peer_opened[q] = true;
    }
    mp->imported = true;
    return 0;
}

int nddwt_mplan_destroy(nddwt_mplan *mp)
{
    release(mp);
    return 0;
}

int nddwt_mplan_world(const nddwt_mplan *mp) { return mp ? mp->world : 0; }
int nddwt_mplan_num_local(const nddwt_mplan *mp) { return mp ? (int)mp->local.size() : 0; }
int nddwt_mplan_is_separable(const nddwt_mplan *mp) { return (mp && mp->separable) ? 1 : 0; }

int nddwt_mplan_slab(const nddwt_mplan *mp, int rank, int64_t *start, int64_t *count)
{
    if (!mp || rank < 0 || rank >= mp->world || !start || !count) { set_error("bad argument"); return NDDWT_ERR_ARG; }
    *start = mp->start[rank];
    *count = mp->count[rank];
    return 0;
}

// host-only routing query (tests; no device work): the planes rank `rank` needs below (which = 0) / above
// (which = 1) its slab for a halo of (below, above) planes, as up to `cap` runs {owner, first local plane,
// count, first halo slot}.  Returns the number of runs.
int nddwt_slab_route(int64_t n_last, int world, int rank, int which, int64_t below, int64_t above, int64_t *runs4, int cap)
{
    if (world < 1 || n_last < world || rank < 0 || rank >= world || !runs4) { set_error("bad argument"); return NDDWT_ERR_ARG; }
    std::vector<int64_t> start, count;
    partition(n_last, world, start, count);
    const std::vector<Run> runs = halo_runs(n_last, start, count, rank, which ? 1 : 0, below, above);
    int k = 0;
    for (const Run &r : runs) {
        if (k >= cap) break;
        runs4[4 * k + 0] = r.owner; runs4[4 * k + 1] = r.idx; runs4[4 * k + 2] = r.cnt; runs4[4 * k + 3] = r.off;
        ++k;
    }
    return (int)runs.size();
}

int nddwt_mplan_set_dilations(nddwt_mplan *mp, const int *dil, int nlevels)
{
    if (!mp || !dil || nlevels < 1 || nlevels > NDDWT_MAX_LEVELS) { set_error("bad dilation list"); return NDDWT_ERR_ARG; }
    if (mp->flags_mode && mp->imported) { set_error("set the dilations before export/import"); return NDDWT_ERR_ARG; }
    int maxdil = 1;
    for (int j = 0; j < nlevels; ++j) {
        if (dil[j] < 1) { set_error("dilation must be >= 1"); return NDDWT_ERR_ARG; }
        maxdil = dil[j] > maxdil ? dil[j] : maxdil;
    }
    int old = 1;
    for (int j = 0; j < NDDWT_MAX_LEVELS; ++j) old = mp->dil[j] > old ? mp->dil[j] : old;
    for (int j = 0; j < nlevels; ++j) mp->dil[j] = dil[j];
    for (RankCtx &c : mp->local) { int rc = nddwt_plan_set_dilations(c.plan, dil, nlevels); if (rc) return rc; }
    mp->separable = true;
    for (RankCtx &c : mp->local) mp->separable = mp->separable && nddwt_plan_is_separable(c.plan) != 0;
    if (maxdil > old) {   // larger halos: rebuild the buffers
        for (RankCtx &c : mp->local) {
            cudaSetDevice(c.device);
            cudaDeviceSynchronize();
            for (int i = 0; i < 2; ++i) { cudaFree(c.approx[i]); cudaFree(c.u_hi[i]); cudaFree(c.over[i]); c.approx[i] = c.u_hi[i] = c.over[i] = nullptr; }
            cudaFree(c.u_lo); cudaFree(c.arena); c.u_lo = nullptr; c.arena = nullptr;
            cudaStreamDestroy(c.cs); cudaStreamDestroy(c.ms); c.cs = c.ms = nullptr;
            for (int i = 0; i < 3; ++i) { cudaStreamDestroy(c.mx[i]); cudaEventDestroy(c.ev_join[i]); c.mx[i] = nullptr; c.ev_join[i] = nullptr; }
            cudaEventDestroy(c.ev_fork); c.ev_fork = nullptr;
            cudaEventDestroy(c.ev_prod); cudaEventDestroy(c.ev_pushed); c.ev_prod = c.ev_pushed = nullptr;
        }
        for (cudaEvent_t e : mp->ev) if (e) cudaEventDestroy(e);
        mp->ev.clear();
        if (mp->status_host) { cudaFreeHost(mp->status_host); mp->status_host = nullptr; }
        return allocate(mp);
    }
    return 0;
}

int nddwt_mplan_set_param(nddwt_mplan *mp, const char *name, int64_t value)
{
    if (!mp || !name) { set_error("null argument"); return NDDWT_ERR_ARG; }
    if (strcmp(name, "z_chunks") == 0) {
        if (value < 0 || value > 64) { set_error("z_chunks must be 0 (automatic) .. 64"); return NDDWT_ERR_ARG; }
        mp->z_chunks = (int)value;
        return 0;
    }
    if (strcmp(name, "comm_streams") == 0) {
        if (value < 1 || value > 4) { set_error("comm_streams must be 1..4"); return NDDWT_ERR_ARG; }
        mp->comm_streams = (int)value;
        return 0;
    }
    for (RankCtx &c : mp->local) { int rc = nddwt_plan_set_param(c.plan, name, value); if (rc) return rc; }
    return 0;
}

int nddwt_mplan_set_shrink(nddwt_mplan *mp, int mode, const double *thr, int nlevels)
{
    if (!mp) { set_error("null plan"); return NDDWT_ERR_ARG; }
    for (RankCtx &c : mp->local) { int rc = nddwt_plan_set_shrink(c.plan, mode, thr, nlevels); if (rc) return rc; }
    return 0;
}

int nddwt_mplan_set_kernel_mode(nddwt_mplan *mp, int mode)
{
    if (!mp) { set_error("null plan"); return NDDWT_ERR_ARG; }
    for (RankCtx &c : mp->local) { int rc = nddwt_plan_set_kernel_mode(c.plan, mode); if (rc) return rc; }
    mp->separable = true;
    for (RankCtx &c : mp->local) mp->separable = mp->separable && nddwt_plan_is_separable(c.plan) != 0;
    return 0;
}

int64_t nddwt_mplan_launch_count(const nddwt_mplan *mp)
{
    if (!mp) return 0;
    int64_t n = mp->extra_launches;
    for (const RankCtx &c : mp->local) n += c.plan->launches;
    return n;
}

int64_t nddwt_mplan_halo_bytes(const nddwt_mplan *mp) { return mp ? mp->copy_bytes : 0; }

int nddwt_mplan_profile(nddwt_mplan *mp, int on)
{
    if (!mp) { set_error("null plan"); return NDDWT_ERR_ARG; }
    for (RankCtx &c : mp->local) { int rc = nddwt_plan_profile(c.plan, on); if (rc) return rc; }
    return 0;
}

int nddwt_mplan_kernel_time(nddwt_mplan *mp, int local_index, int kind, double *total_ms, int64_t *count)
{
    if (!mp || local_index < 0 || local_index >= (int)mp->local.size()) { set_error("bad argument"); return NDDWT_ERR_ARG; }
    return nddwt_plan_kernel_time(mp->local[local_index].plan, kind, total_ms, count);
}
int nddwt_mplan_wait_timeouts(const nddwt_mplan *mp) { return (mp && mp->status_host) ? (int)*(volatile uint32_t *)mp->status_host : 0; }

int nddwt_mplan_sync(nddwt_mplan *mp)
{
    if (!mp) { set_error("null plan"); return NDDWT_ERR_ARG; }
    for (RankCtx &c : mp->local) {
        NDDWT_CUDA(cudaSetDevice(c.device));
        NDDWT_CUDA(cudaStreamSynchronize(c.ms));
        for (int i = 0; i < 3; ++i) NDDWT_CUDA(cudaStreamSynchronize(c.mx[i]));
        NDDWT_CUDA(cudaStreamSynchronize(c.cs));
    }
    if (nddwt_mplan_wait_timeouts(mp) > 0) { set_error("a peer flag wait timed out (a rank did not take part in the call?)"); return NDDWT_ERR_CUDA; }
    return 0;
}

static int check_call(nddwt_mplan *mp, const void *const *a, const void *const *b, int level)
{
    if (!mp || !a || !b) { set_error("null argument"); return NDDWT_ERR_ARG; }
    if (level < 1 || level > NDDWT_MAX_LEVELS) { set_error("level must be in 1..16"); return NDDWT_ERR_ARG; }
    if (!mp->imported) { set_error("rank plan: exchange the export blobs and call nddwt_mplan_import first"); return NDDWT_ERR_ARG; }
    for (size_t i = 0; i < mp->local.size(); ++i)
        if (!a[i] || !b[i]) { set_error("null slab pointer"); return NDDWT_ERR_ARG; }
    return 0;
}

int nddwt_mplan_dec(nddwt_mplan *mp, const void *const *x_slabs, void *const *coeff_slabs, int level, void *const *streams)
{
    int rc = check_call(mp, x_slabs, (const void *const *)coeff_slabs, level);
    if (rc) return rc;
    const int nl = (int)mp->local.size(), nd = 1 << mp->ndims, L = mp->L_last;
    std::vector<cudaStream_t> cs(nl);
    for (int i = 0; i < nl; ++i) cs[i] = streams ? reinterpret_cast<cudaStream_t>(streams[i]) : mp->local[i].cs;   // a NULL entry is the legacy default stream
    if (mp->world == 1) return nddwt_dec(mp->local[0].plan, x_slabs[0], coeff_slabs[0], level, cs[0]);
    if (mp->separable) {
        const Chunks ch = make_chunks(mp);
        if (ch.C > 1) return dec_chunked(mp, x_slabs, coeff_slabs, level, cs, ch);
    }
    std::vector<const void *> a_in(x_slabs, x_slabs + nl);
    for (int j = 1; j <= level; ++j) {
        const int par = j & 1, dil = mp->dil[j - 1];
        const int64_t below = (int64_t)(L / 2 - 1) * dil, above = (int64_t)(L / 2) * dil;
        // ---- every rank pushes the planes of a_{j-1} its neighbours need (comm stream, copy engines)
        for (int i = 0; i < nl; ++i) {
            RankCtx &c = mp->local[i];
            NDDWT_CUDA(cudaSetDevice(c.device));
            if (j == 1) NDDWT_CUDA(cudaEventRecord(c.ev_prod, cs[i]));      // x is ready once the caller's stream gets here
            NDDWT_CUDA(cudaStreamWaitEvent(c.ms, c.ev_prod, 0));
            PushSrc src = {reinterpret_cast<const char *>(a_in[i]), {0, 0}};
            rc = push_halos(mp, c, &src, 1, below, above, par, CH_HALO_FILLED0 + par, CH_HALO_FREE0 + par);
            if (rc) return rc;
        }
        // ---- last-dim pass (the only reader of the halos), then the tile pass that yields a_j
        std::vector<std::vector<void *>> bands(nl, std::vector<void *>(nd));
        for (int i = 0; i < nl; ++i) {
            RankCtx &c = mp->local[i];
            band_ptrs(mp, c, reinterpret_cast<char *>(coeff_slabs[i]), level, j, bands[i].data());
            bands[i][0] = (j == level) ? coeff_slabs[i] : c.approx[j & 1];
            rc = wait_halos(mp, c, cs[i], below, above, CH_HALO_FILLED0 + par);
            if (rc) return rc;
            const void *hl = inbox_ptr(mp, c.rank, par, 0), *hh = inbox_ptr(mp, c.rank, par, 1);
            rc = nddwt_dec_level_slab_part(c.plan, j, 1, a_in[i], hl, hh, bands[i].data(), cs[i]);
            if (rc) return rc;
            rc = free_halos(mp, c, cs[i], below, above, CH_HALO_FREE0 + par);
            if (rc) return rc;
            rc = nddwt_dec_level_slab_part(c.plan, j, 2, a_in[i], hl, hh, bands[i].data(), cs[i]);
            if (rc) return rc;
            NDDWT_CUDA(cudaSetDevice(c.device));
            NDDWT_CUDA(cudaEventRecord(c.ev_prod, cs[i]));                  // a_j complete: the next push may start ...
        }
        for (int i = 0; i < nl; ++i) {                                      // ... while the remaining detail bands are computed
            RankCtx &c = mp->local[i];
            rc = nddwt_dec_level_slab_part(c.plan, j, 3, a_in[i], nullptr, nullptr, bands[i].data(), cs[i]);
            if (rc) return rc;
            a_in[i] = bands[i][0];
        }
    }
    return 0;
}

int nddwt_mplan_rec(nddwt_mplan *mp, const void *const *coeff_slabs, void *const *x_slabs, int level, void *const *streams)
{
    int rc = check_call(mp, coeff_slabs, (const void *const *)x_slabs, level);
    if (rc) return rc;
    const int nl = (int)mp->local.size(), nd = 1 << mp->ndims, L = mp->L_last;
    std::vector<cudaStream_t> cs(nl);
    for (int i = 0; i < nl; ++i) cs[i] = streams ? reinterpret_cast<cudaStream_t>(streams[i]) : mp->local[i].cs;   // a NULL entry is the legacy default stream
    if (mp->world == 1) return nddwt_rec(mp->local[0].plan, coeff_slabs[0], x_slabs[0], level, cs[0]);
    if (mp->separable) {
        const Chunks ch = make_chunks(mp);
        if (ch.C > 1) return rec_chunked(mp, coeff_slabs, x_slabs, level, cs, ch);
    }
    std::vector<const void *> a(coeff_slabs, coeff_slabs + nl);     // slot 0 = deepest approximation
    std::vector<std::vector<void *>> bands(nl, std::vector<void *>(nd));
    auto set_bands = [&](int i, int j, const void *approx) {
        band_ptrs(mp, mp->local[i], reinterpret_cast<char *>(const_cast<void *>(coeff_slabs[i])), level, j, bands[i].data());
        bands[i][0] = const_cast<void *>(approx);
    };
    if (mp->separable) {
        // scatter form, overlapped: detail-only half of stage 1 of the NEXT level runs while the overhangs move
        for (int i = 0; i < nl; ++i) {
            RankCtx &c = mp->local[i];
            set_bands(i, level, a[i]);
            rc = nddwt_rec_level_slab_stage1_part(c.plan, level, 2, bands[i].data(), c.u_lo, c.u_hi[level & 1], cs[i]);
            if (rc) return rc;
        }
        for (int j = level; j >= 1; --j) {
            const int k = j & 1, dil = mp->dil[j - 1];
            const int64_t below = (int64_t)(L / 2 - 1) * dil, above = (int64_t)(L / 2) * dil;
            std::vector<void *> dst(nl);
            for (int i = 0; i < nl; ++i) {
                RankCtx &c = mp->local[i];
                dst[i] = (j == 1) ? x_slabs[i] : c.approx[j & 1];
                set_bands(i, j, a[i]);
                rc = nddwt_rec_level_slab_stage1_part(c.plan, j, 1, bands[i].data(), c.u_lo, c.u_hi[k], cs[i]);
                if (rc) return rc;
                NDDWT_CUDA(cudaSetDevice(c.device));
                if (c.pushed_once) NDDWT_CUDA(cudaStreamWaitEvent(cs[i], c.ev_pushed, 0));   // previous overhangs have left
                rc = nddwt_rec_level_slab_stage2_scatter(c.plan, j, c.u_lo, c.u_hi[k], dst[i], c.over[0], c.over[1], cs[i]);
                if (rc) return rc;
                NDDWT_CUDA(cudaEventRecord(c.ev_prod, cs[i]));
                NDDWT_CUDA(cudaStreamWaitEvent(c.ms, c.ev_prod, 0));
                // push my overhang planes to their owners' stage buffers
                LaunchTimer lt(c.plan, KIND_COMM, c.ms);
                for (int dq = 1; dq < mp->world; ++dq) {
                    const int q = (c.rank + dq) % mp->world;
                    bool any = false;
                    for (const Piece &p : pieces_into(mp, q, below, above)) {
                        if (p.src != c.rank) continue;
                        if (!any) {
                            uint32_t &rounds = c.rounds[(size_t)q * NCH + CH_STAGE_FILLED];
                            if (rounds > 0) { rc = fab_wait(mp, c, q, CH_STAGE_FREE, c.ms, rounds); if (rc) return rc; }
                            ++rounds;
                            NDDWT_CUDA(cudaSetDevice(c.device));
                            rc = fork_streams(mp, c);
                            if (rc) return rc;
                            any = true;
                        }
                        char *d = mp->peer_arena[q] + mp->off_stage + (size_t)p.stage_off * mp->plane_bytes;
                        const char *s = reinterpret_cast<const char *>(c.over[p.which]) + (size_t)p.over_off * mp->plane_bytes;
                        rc = peer_copy(mp, c, d, s, (size_t)p.cnt * mp->plane_bytes);
                        if (rc) return rc;
                    }
                    if (any) {
                        rc = join_streams(mp, c);
                        if (rc) return rc;
                        rc = fab_signal(mp, c, q, CH_STAGE_FILLED, c.ms);
                        if (rc) return rc;
                    }
                }
                NDDWT_CUDA(cudaSetDevice(c.device));
                NDDWT_CUDA(cudaEventRecord(c.ev_pushed, c.ms));
                c.pushed_once = true;
            }
            if (j > 1)
                for (int i = 0; i < nl; ++i) {
                    RankCtx &c = mp->local[i];
                    set_bands(i, j - 1, coeff_slabs[i]);   // detail bands only: bands[0] is not read by part 2
                    rc = nddwt_rec_level_slab_stage1_part(c.plan, j - 1, 2, bands[i].data(), c.u_lo, c.u_hi[(j - 1) & 1], cs[i]);
                    if (rc) return rc;
                }
            for (int i = 0; i < nl; ++i) {
                RankCtx &c = mp->local[i];
                const std::vector<Piece> pieces = pieces_into(mp, c.rank, below, above);
                std::vector<char> from(mp->world, 0);
                for (const Piece &p : pieces) if (p.src != c.rank) from[p.src] = 1;
                for (int q = 0; q < mp->world; ++q)
                    if (from[q]) { rc = fab_wait(mp, c, q, CH_STAGE_FILLED, cs[i]); if (rc) return rc; }
                // every local plane is read and written once, whatever number of overhangs lands on it
                AccParams acc;
                acc.n = 0;
                bool fits = true;
                std::vector<int> slot((size_t)c.count, -1);
                for (const Piece &p : pieces) {
                    const char *s = (p.src == c.rank)
                                        ? reinterpret_cast<const char *>(c.over[p.which]) + (size_t)p.over_off * mp->plane_bytes
                                        : c.arena + mp->off_stage + (size_t)p.stage_off * mp->plane_bytes;
                    for (int64_t t = 0; t < p.cnt && fits; ++t) {
                        int &k = slot[p.idx + t];
                        if (k < 0) {
                            if (acc.n >= ACC_MAXP) { fits = false; break; }
                            k = acc.n++;
                            acc.item[k].dst = reinterpret_cast<char *>(dst[i]) + (size_t)(p.idx + t) * mp->plane_bytes;
                            acc.item[k].ns = 0;
                        }
                        if (acc.item[k].ns >= ACC_MAXS) { fits = false; break; }
                        acc.item[k].src[acc.item[k].ns++] = s + (size_t)t * mp->plane_bytes;
                    }
                }
                if (fits) {
                    NDDWT_CUDA(cudaSetDevice(c.device));
                    rc = accumulate_planes(c.plan, acc, mp->plane_elems, cs[i]);
                    if (rc) return rc;
                } else {
                    for (const Piece &p : pieces) {
                        char *d = reinterpret_cast<char *>(dst[i]) + (size_t)p.idx * mp->plane_bytes;
                        const char *s = (p.src == c.rank)
                                            ? reinterpret_cast<const char *>(c.over[p.which]) + (size_t)p.over_off * mp->plane_bytes
                                            : c.arena + mp->off_stage + (size_t)p.stage_off * mp->plane_bytes;
                        rc = nddwt_accumulate(c.plan, d, s, p.cnt * mp->plane_elems, cs[i]);
                        if (rc) return rc;
                    }
                }
                for (int q = 0; q < mp->world; ++q)
                    if (from[q]) { rc = fab_signal(mp, c, q, CH_STAGE_FREE, cs[i]); if (rc) return rc; }
                a[i] = dst[i];
            }
        }
        return 0;
    }
    // gather form (plans without the separable fused 4-D path): stage 1 locally, push the halos of u_lo and
    // u_hi (L/2 below, L/2-1 above), stage 2
    for (int j = level; j >= 1; --j) {
        const int dil = mp->dil[j - 1], par = j & 1;
        const int64_t below = (int64_t)(L / 2) * dil, above = (int64_t)(L / 2 - 1) * dil;
        std::vector<void *> dst(nl);
        for (int i = 0; i < nl; ++i) {
            RankCtx &c = mp->local[i];
            dst[i] = (j == 1) ? x_slabs[i] : c.approx[j & 1];
            set_bands(i, j, a[i]);
            rc = nddwt_rec_level_slab_stage1(c.plan, j, bands[i].data(), c.u_lo, c.u_hi[0], cs[i]);
            if (rc) return rc;
            NDDWT_CUDA(cudaSetDevice(c.device));
            NDDWT_CUDA(cudaEventRecord(c.ev_prod, cs[i]));
            NDDWT_CUDA(cudaStreamWaitEvent(c.ms, c.ev_prod, 0));
            // inbox layout of nddwt_rec_level_slab_stage2: u_lo planes, then u_hi planes, on each side
            PushSrc srcs[2] = {{reinterpret_cast<const char *>(c.u_lo), {0, 0}},
                               {reinterpret_cast<const char *>(c.u_hi[0]), {below, above}}};
            rc = push_halos(mp, c, srcs, 2, below, above, par, CH_RECH_FILLED, CH_RECH_FREE);
            if (rc) return rc;
        }
        for (int i = 0; i < nl; ++i) {
            RankCtx &c = mp->local[i];
            rc = wait_halos(mp, c, cs[i], below, above, CH_RECH_FILLED);
            if (rc) return rc;
            rc = nddwt_rec_level_slab_stage2(c.plan, j, c.u_lo, c.u_hi[0], inbox_ptr(mp, c.rank, par, 0),
                                             inbox_ptr(mp, c.rank, par, 1), dst[i], cs[i]);
            if (rc) return rc;
            rc = free_halos(mp, c, cs[i], below, above, CH_RECH_FREE);
            if (rc) return rc;
            a[i] = dst[i];
        }
    }
    return 0;
}


// ---- host arrays in, host arrays out (the shape of nd_dwt_mex for a MATLAB / C caller that owns all GPUs):
// the slabs of a column-major array along its LAST dimension are contiguous, so rank r's part of x is one
// block of the host array and its part of every band one block of that band; every GPU moves its own slabs
// over its own PCIe link.  Needs a one-process plan (nddwt_mplan_create).
static int host_staging(nddwt_mplan *mp, int level)
{
    const size_t nb = (size_t)nddwt_num_bands(mp->ndims, level);
    for (RankCtx &c : mp->local) {
        NDDWT_CUDA(cudaSetDevice(c.device));
        const size_t slab = (size_t)c.count * mp->plane_bytes;
        if (!c.host_x) NDDWT_CUDA(cudaMalloc(&c.host_x, slab));
        if (c.host_c_bytes < slab * nb) {
            if (c.host_c) { cudaFree(c.host_c); c.host_c = nullptr; c.host_c_bytes = 0; }
            NDDWT_CUDA(cudaMalloc(&c.host_c, slab * nb));
            c.host_c_bytes = slab * nb;
        }
    }
    return 0;
}

int nddwt_mplan_dec_host(nddwt_mplan *mp, const void *x_host, void *coeffs_host, int level)
{
    if (!mp || !x_host || !coeffs_host) { set_error("null argument"); return NDDWT_ERR_ARG; }
    if ((int)mp->local.size() != mp->world) { set_error("host entry points need a one-process plan (nddwt_mplan_create)"); return NDDWT_ERR_ARG; }
    if (level < 1 || level > NDDWT_MAX_LEVELS) { set_error("level must be in 1..16"); return NDDWT_ERR_ARG; }
    int rc = host_staging(mp, level);
    if (rc) return rc;
    const size_t nb = (size_t)nddwt_num_bands(mp->ndims, level);
    const size_t band_bytes = (size_t)mp->dims[mp->ndims - 1] * mp->plane_bytes;
    std::vector<const void *> xs;
    std::vector<void *> cs;
    for (RankCtx &c : mp->local) {
        NDDWT_CUDA(cudaSetDevice(c.device));
        NDDWT_CUDA(cudaMemcpyAsync(c.host_x, reinterpret_cast<const char *>(x_host) + (size_t)c.start * mp->plane_bytes,
                                   (size_t)c.count * mp->plane_bytes, cudaMemcpyHostToDevice, c.cs));
        xs.push_back(c.host_x);
        cs.push_back(c.host_c);
    }
    rc = nddwt_mplan_dec(mp, xs.data(), cs.data(), level, nullptr);
    if (rc) return rc;
    for (RankCtx &c : mp->local) {
        NDDWT_CUDA(cudaSetDevice(c.device));
        const size_t slab = (size_t)c.count * mp->plane_bytes;
        for (size_t b = 0; b < nb; ++b)
            NDDWT_CUDA(cudaMemcpyAsync(reinterpret_cast<char *>(coeffs_host) + b * band_bytes + (size_t)c.start * mp->plane_bytes,
                                       reinterpret_cast<const char *>(c.host_c) + b * slab, slab, cudaMemcpyDeviceToHost, c.cs));
    }
    return nddwt_mplan_sync(mp);
}

int nddwt_mplan_rec_host(nddwt_mplan *mp, const void *coeffs_host, void *x_host, int level)
{
    if (!mp || !x_host || !coeffs_host) { set_error("null argument"); return NDDWT_ERR_ARG; }
    if ((int)mp->local.size() != mp->world) { set_error("host entry points need a one-process plan (nddwt_mplan_create)"); return NDDWT_ERR_ARG; }
    if (level < 1 || level > NDDWT_MAX_LEVELS) { set_error("level must be in 1..16"); return NDDWT_ERR_ARG; }
    int rc = host_staging(mp, level);
    if (rc) return rc;
    const size_t nb = (size_t)nddwt_num_bands(mp->ndims, level);
    const size_t band_bytes = (size_t)mp->dims[mp->ndims - 1] * mp->plane_bytes;
    std::vector<const void *> cs;
    std::vector<void *> xs;
    for (RankCtx &c : mp->local) {
        NDDWT_CUDA(cudaSetDevice(c.device));
        const size_t slab = (size_t)c.count * mp->plane_bytes;
        for (size_t b = 0; b < nb; ++b)
            NDDWT_CUDA(cudaMemcpyAsync(reinterpret_cast<char *>(c.host_c) + b * slab,
                                       reinterpret_cast<const char *>(coeffs_host) + b * band_bytes + (size_t)c.start * mp->plane_bytes,
                                       slab, cudaMemcpyHostToDevice, c.cs));
        cs.push_back(c.host_c);
        xs.push_back(c.host_x);
    }
    rc = nddwt_mplan_rec(mp, cs.data(), xs.data(), level, nullptr);
    if (rc) return rc;
    for (RankCtx &c : mp->local) {
        NDDWT_CUDA(cudaSetDevice(c.device));
        NDDWT_CUDA(cudaMemcpyAsync(reinterpret_cast<char *>(x_host) + (size_t)c.start * mp->plane_bytes, c.host_x,
                                   (size_t)c.count * mp->plane_bytes, cudaMemcpyDeviceToHost, c.cs));
    }
    return nddwt_mplan_sync(mp);
}

}  // extern "C"
