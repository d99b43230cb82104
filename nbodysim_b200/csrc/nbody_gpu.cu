// nbody_gpu.cu -- the C ABI (include/nbody_gpu.h) and the per-step driver.
//
// Replaces, behind a C boundary, the body of Simulation::step()'s hot path
// (Simulation.hpp:67-75 -> iterate :116-164 -> attract :176-214 -> Quadtree::acc, Quadtree.hpp:113-155):
// per step, force accumulation on the current positions, then kick-drift.  The reference's
// std::async fan-out over target chunks (Simulation.hpp:190-213) becomes (a) a grid of CTAs per GPU
// and (b) a shard of targets per GPU with an NCCL allgather of the new positions each step,
// overlapped with the force pass over the locally-owned sources.
#include "../../include/nbody_gpu.h"
#include "kernels.h"
#include "nccl_dyn.h"
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <new>
#include <thread>
#include <vector>

using namespace nb;

namespace {

thread_local char g_init_err[512] = "";

struct Range {          // one force launch of the per-step plan
    int j_blk0, j_nblk; // source blocks
    int splits, slot0;
    bool remote;        // needs the allgather of the previous step to have landed
    int sk_ctas = 0;    // > 0: stream-K launch on that many persistent CTAs (fast fp32 kernel); `splits` then holds the slots it needs
};

struct Dev {
    int device = 0;
    int rank = 0;                 // global rank of this GPU
    cudaStream_t stream = nullptr, comm_stream = nullptr;
    bool own_stream = true;
    cudaEvent_t ev_pushed[2] = {nullptr, nullptr};   // P2P exchange: this GPU's integrate-and-push of step parity 0/1 is done
    cudaEvent_t ev_integrated = nullptr, ev_gathered = nullptr, ev_t[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    void *posm[2] = {nullptr, nullptr};
    void *vel = nullptr, *acc = nullptr, *accp = nullptr, *aos = nullptr;
    double *energy5 = nullptr;
    unsigned long long *walk_visits = nullptr;          // profiled Barnes-Hut steps: node records visited by the walk
    unsigned *h_status = nullptr, *d_status = nullptr;   // pinned + mapped: [0] Barnes-Hut cell overflow, [1] collision buffer overflow
    ncclComm_t comm = nullptr;
    // cross-process exchange (one process per GPU): the peers' position buffers and flag arrays, mapped by CUDA IPC
    void *peer_posm[NBODY_MAX_GPUS][2] = {};
    unsigned long long *peer_flags[NBODY_MAX_GPUS] = {};
    unsigned long long *flags = nullptr;     // own flag array: flags[r] = steps of rank r whose pushes have landed here
    unsigned *ipc_words = nullptr;           // [0] CTA arrival counter, [1] time-out status, [4..] barrier scratch
    size_t shard_start = 0, shard_count = 0; // bodies (padded index space)
    int cur = 0;
    bool gathered_pending = false;           // an allgather into posm[cur] may still be in flight
    std::vector<Range> plan;
    int nslots = 0;
    bool fused = false;
    bool small_tile = false;      // fast kernel runs the 512-target tile geometry (small shards)
    int force_ctas = 0;
    BhWorkspace bh;               // Barnes-Hut path only
    CollideWorkspace col;         // collision pass only
    // CUDA graph of two consecutive steps (buffer parity returns to the start), keyed on dt and parity
    cudaGraphExec_t graph = nullptr;
    float graph_dt = 0.f;
    int graph_cur = -1;
    unsigned long long graph_launches = 0, graph_interactions = 0;
    // ... and of ONE step per buffer parity, for callers that step once per call (a viewer, the e2e loop)
    cudaGraphExec_t graph1[2] = {nullptr, nullptr};
    float graph1_dt[2] = {0.f, 0.f};
    unsigned long long graph1_launches = 0, graph1_interactions = 0;
};

} // namespace

struct nbody_ctx {
    nbody_params p;
    size_t n = 0, n_padded = 0;
    int world = 1;               // total GPUs
    bool f64 = false;
    bool small_tile = false;     // fast kernel geometry: 512-target tiles (small shards) instead of 2048
    bool p2p = false;            // positions exchanged by peer stores from the integrator kernel (one process, ngpus > 1)
    bool ipc = false;            // same exchange across processes: peers mapped by CUDA IPC, completion flags in peer memory
    unsigned long long peer_timeout_ns = 120ull * 1000000000ull;
    unsigned long long step_index = 0;
    bool bh = false;             // force_algo == NBODY_FORCE_BARNES_HUT
    bool streamk = true;         // fast fp32 kernel in stream-K form unless j_splits > 0 (tuning: NBODY_FORCE_FORM=split)
    bool uniform = false;        // every massive body has the same mass: 11-op force kernel
    float uniform_mass = 0.f;
    size_t esz = 4;
    std::vector<Dev> devs;
    nbody_body_t *h_stage = nullptr; // pinned, n records (download merges / uploads)
    int sm_count = 0, sm_clock_khz = 0, ctas_per_sm = 0, ctas_per_sm_small = 0, sk_ctas_per_sm = 0, sk_ctas_per_sm_small = 0;
    unsigned long long launches = 0, interactions = 0;
    int profile_next = 0;
    unsigned long long short_calls = 0;  // nbody_gpu_step calls of fewer than 8 steps so far
    float last_force_ms = 0.f, last_integ_ms = 0.f, last_build_ms = 0.f, last_collide_ms = 0.f;
    unsigned long long last_visits = 0, last_visits_max = 0;
    char err[512];
    nbody_ctx() { err[0] = 0; }
};

namespace {

void set_err(nbody_ctx *c, const char *fmt, ...)
{
    char *dst = c ? c->err : g_init_err;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(dst, 512, fmt, ap);
    va_end(ap);
}

#define CU(call)                                                                               \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            set_err(ctx, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return e_ == cudaErrorMemoryAllocation ? NBODY_ENOMEM : NBODY_ECUDA;               \
        }                                                                                      \
    } while (0)
#define NC(call)                                                                               \
    do {                                                                                       \
        int r_ = (call);                                                                       \
        if (r_ != 0) {                                                                         \
            set_err(ctx, "%s:%d %s -> %s", __FILE__, __LINE__, #call, nccl().GetErrorString(r_)); \
            return NBODY_ENCCL;                                                                \
        }                                                                                      \
    } while (0)

IntegParams make_ip(const nbody_ctx *c, float dt)
{
    IntegParams ip;
    ip.dt = dt;
    ip.G = c->p.G;
    ip.flags = c->p.integ_flags;
    ip.max_velocity = c->p.max_velocity;
    ip.max_velocity_sq = c->p.max_velocity * c->p.max_velocity;
    ip.soft_boundary = c->p.boundary_radius * c->p.soft_boundary;
    ip.soft_boundary_sq = ip.soft_boundary * ip.soft_boundary;
    ip.boundary_force = c->p.boundary_force;
    ip.damping = c->p.damping;
    return ip;
}

// Choose the number of source-range splits so that the CTAs of one launch fill whole waves of
// the GPU (slots = SMs x resident CTAs per SM): wave-quantisation efficiency
//   eff(S) = units / (ceil(units/slots) * slots),  units = target_tiles * S.
// Prefer the smallest S within 1 % (relative) of the best; each chunk keeps >= min_chunk_blks source blocks.
int choose_splits(int tiles, int j_nblk, int slots, int min_chunk_blks, int max_splits)
{
    if (tiles <= 0 || slots <= 0) return 1;
    int smax = std::max(1, std::min(max_splits, j_nblk / std::max(1, min_chunk_blks)));
    double best = -1.0;
    int best_s = 1;
    for (int s = 1; s <= smax; ++s) {
        const long long units = (long long)tiles * s;
        const long long waves = (units + slots - 1) / slots;
        double eff = (double)units / (double)(waves * slots);
        // a launch of very few waves also suffers the imbalance of a dynamic tail: favour >= 4 waves
        if (waves < 4) eff *= 0.97;
        if (eff > best * 1.01) { best = eff; best_s = s; }   // prefer the smallest S within 1 % of the best
    }
    return best_s;
}

int plan_device(nbody_ctx *c, Dev &d)
{
    const int nblk = (int)(c->n_padded / BLK);
    const int ib0 = (int)(d.shard_start / BLK), ibn = (int)(d.shard_count / BLK);
    const bool refc = !c->f64 && c->p.rsqrt_mode == NBODY_RSQRT_REFCOMPAT;
    d.plan.clear();
    d.fused = false;
    d.small_tile = false;
    int tiles, slots, min_chunk;
    if (c->f64) { tiles = ibn / F64_TILE_BLKS; slots = c->sm_count * 4; min_chunk = 4; }
    else if (refc) { tiles = ibn * 2; slots = c->sm_count * 4; min_chunk = 1; }
    else {
        tiles = ibn / FAST_TILE_BLKS; slots = c->sm_count * std::max(1, c->ctas_per_sm); min_chunk = 8;
        d.small_tile = c->small_tile;                     // decided once in nbody_gpu_init (it fixes the padding)
        if (d.small_tile) { tiles = ibn / SMALL_TILE_BLKS; slots = c->sm_count * std::max(1, c->ctas_per_sm_small); min_chunk = 1; }
    }

    struct Seg { int b0, nb; bool remote; };
    std::vector<Seg> segs;
    if (c->bh) { // one tree over all sources, one walk launch, one partial slot
        d.plan.push_back(Range{0, nblk, 1, 0, c->world > 1});
        d.nslots = 1;
        d.force_ctas = (int)((c->n + 127) / 128);
        return NBODY_OK;
    }
    if (c->world == 1 || refc) {
        segs.push_back({0, nblk, c->world > 1});
    } else {
        segs.push_back({ib0, ibn, false});                        // locally owned sources first
        if (ib0 > 0) segs.push_back({0, ib0, true});
        if (ib0 + ibn < nblk) segs.push_back({ib0 + ibn, nblk - ib0 - ibn, true});
    }
    int slot = 0;
    const bool fastk = !c->f64 && !refc;
    for (const Seg &s : segs) {
        int S = 1, sk = 0;
        if (fastk && c->p.j_splits == 0 && c->p.fuse_integrator != 1 && c->streamk) {
            // stream-K: one persistent CTA per SM slot, equal runs of (tile, source stage) units; a tile's partial sums land
            // in as many slots as CTAs share it (two at N = 1M) instead of one per source split (13 there)
            const int per_sm = d.small_tile ? c->sk_ctas_per_sm_small : c->sk_ctas_per_sm;
            const int stage_blks = d.small_tile ? SMALL_STAGE_BLKS : FAST_STAGE_BLKS;
            const int stages = (s.nb + stage_blks - 1) / stage_blks;
            // never more CTAs than units: every CTA then owns a non-empty run, so the CTAs that share a tile are consecutive
            // and every slot the integrator reads has been written
            sk = (int)std::max<long long>(1, std::min<long long>((long long)c->sm_count * std::max(1, per_sm), (long long)tiles * stages));
            S = force_f32_streamk_slots(tiles, stages, sk);
        } else if (!refc) {
            if (c->p.j_splits > 0) S = std::min(c->p.j_splits, s.nb);
            else S = choose_splits(tiles, s.nb, slots, min_chunk, 64);
        }
        Range r{s.b0, s.nb, S, slot, s.remote};
        r.sk_ctas = sk;
        d.plan.push_back(r);
        slot += S;
    }
    d.nslots = slot;
    if (!c->f64 && !refc && !d.small_tile && c->world == 1 && d.plan.size() == 1 && d.plan[0].splits == 1 && d.plan[0].sk_ctas == 0) {
        d.fused = (c->p.fuse_integrator != 0);
    }
    if (c->p.fuse_integrator == 1 && !(d.plan.size() == 1 && d.plan[0].splits == 1 && !c->f64 && !refc && !d.small_tile)) {
        // explicit request that cannot be honoured with this plan: fall back to separate kernels
        d.fused = false;
    }
    d.force_ctas = 0;
    for (const Range &r : d.plan) d.force_ctas += r.sk_ctas > 0 ? r.sk_ctas : tiles * r.splits;
    return NBODY_OK;
}

int alloc_device(nbody_ctx *ctx, Dev &d)
{
    CU(cudaSetDevice(d.device));
    const size_t esz = ctx->esz;
    const size_t full = ctx->n_padded * 4 * esz, shard = d.shard_count * 4 * esz;
    if (d.own_stream) CU(cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking));
    if (ctx->world > 1 && !ctx->p2p) CU(cudaStreamCreateWithFlags(&d.comm_stream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&d.ev_integrated, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&d.ev_gathered, cudaEventDisableTiming));
    for (int k = 0; k < 2; ++k) CU(cudaEventCreateWithFlags(&d.ev_pushed[k], cudaEventDisableTiming));
    for (int k = 0; k < 6; ++k) CU(cudaEventCreate(&d.ev_t[k]));
    CU(cudaMalloc(&d.posm[0], full));
    CU(cudaMalloc(&d.posm[1], full));
    CU(cudaMalloc(&d.vel, shard));
    CU(cudaMalloc(&d.acc, shard));
    CU(cudaMalloc(&d.accp, shard * (size_t)std::max(1, d.nslots)));
    CU(cudaMalloc(&d.aos, ctx->n_padded * sizeof(nbody_body_t)));
    CU(cudaMalloc(&d.energy5, 5 * sizeof(double)));
    CU(cudaMalloc(&d.walk_visits, 2 * sizeof(unsigned long long)));     // total, longest walk
    // sticky overflow flags the kernels raise and sync / download / energy check: host memory mapped into the device
    // address space, so checking them costs no copy
    CU(cudaHostAlloc((void **)&d.h_status, 64, cudaHostAllocMapped));
    memset(d.h_status, 0, 64);
    CU(cudaHostGetDevicePointer((void **)&d.d_status, d.h_status, 0));
    size_t cl_max = 65536;                      // scenes up to this size run their build / collision pass as one cluster kernel (sort_impl = 0)
    if (const char *cm = getenv("NBODY_CLUSTER_MAX_N")) cl_max = (size_t)std::max(0ll, atoll(cm));
    if (ctx->p.collide) {
        CU(d.col.alloc(ctx->n, ctx->p.sort_impl, cl_max));
        d.col.status = d.d_status;
    }
    if (ctx->bh) {
        double factor = 4.0;
        if (const char *nf = getenv("NBODY_BH_NODE_FACTOR")) factor = atof(nf);
        {
            // measured (profiles/r2_cluster_build_trace.txt): the single-cluster BUILD is slower than one kernel per phase on
            // the whole GPU (267 vs 157 us at n = 25,000), so it runs only on request; the single-cluster COLLISION PASS is
            // faster (60 vs ~300 us cold) and is the default for small scenes
            const cudaError_t be = d.bh.alloc(ctx->n, ctx->p.dims, factor, ctx->p.sort_impl == 2 ? 2 : 1, cl_max);
            if (be == cudaErrorNotSupported) {
                set_err(ctx, "sort_impl = 2 (single-cluster build) needs a device that can host a cluster of >= 8 CTAs and n <= CTAs x 49152");
                return NBODY_EINVAL;
            }
            CU(be);
        }
        d.bh.status = d.d_status;
        d.bh.warp_walk = ctx->p.bh_walk == 2 || (ctx->p.bh_walk == 0 && (ctx->p.dims == 3 || ctx->p.theta < 0.7f));
        if (const char *ww = getenv("NBODY_BH_WALK_WINDOW")) d.bh.walk_window = (unsigned)std::max(1, atoi(ww));   // tuning override
    }
    return NBODY_OK;
}

int ipc_barrier(nbody_ctx *ctx);


// Initial upload (`shard_only` = false): every local GPU receives and packs all n bodies.
// Re-upload of a distributed context (`shard_only` = true, one process per GPU): only this rank's records cross
// PCIe; the packed shard then reaches the other ranks over NVLink -- copies into the peers' mapped buffers (CUDA
// IPC exchange) or one in-place ncclAllGather -- bracketed by stream-ordered barriers so that no rank's buffer is
// written while a peer still reads it, and every rank's copies have landed before anyone proceeds.
int upload_state(nbody_ctx *ctx, const nbody_body_t *bodies, bool shard_only, bool check_masses)
{
    // the uniform-mass kernel stays valid only while the masses stay as first uploaded: the pack kernel compares
    const float check_mass = (ctx->uniform && check_masses) ? ctx->uniform_mass : 0.f;
    if (shard_only && ctx->ipc) {
        int rc = ipc_barrier(ctx);
        if (rc != NBODY_OK) return rc;
    }
    for (Dev &d : ctx->devs) {
        CU(cudaSetDevice(d.device));
        d.cur = 0;
        d.gathered_pending = false;
        const size_t full_bytes = ctx->n_padded * 4 * ctx->esz;
        if (!shard_only) {
            CU(cudaMemcpyAsync(d.aos, bodies, ctx->n * sizeof(nbody_body_t), cudaMemcpyHostToDevice, d.stream));
            CU(launch_pack(d.aos, 0, 0, ctx->n_padded, ctx->n, d.shard_start, d.shard_count, d.posm[0], d.vel,
                           d.acc, ctx->f64, ctx->p.dims, check_mass, d.d_status, d.stream));
            ctx->launches++;
            // padding must also be valid in the other buffer (masses are copied by the integrator only
            // for the shard's own blocks, so seed both buffers with the full packed state)
            CU(cudaMemcpyAsync(d.posm[1], d.posm[0], full_bytes, cudaMemcpyDeviceToDevice, d.stream));
            continue;
        }
        const size_t real = d.shard_start < ctx->n ? std::min(ctx->n - d.shard_start, d.shard_count) : 0;
        if (real) CU(cudaMemcpyAsync(d.aos, bodies + d.shard_start, real * sizeof(nbody_body_t), cudaMemcpyHostToDevice, d.stream));
        CU(launch_pack(d.aos, d.shard_start, d.shard_start, d.shard_count, ctx->n, d.shard_start, d.shard_count, d.posm[0], d.vel,
                       d.acc, ctx->f64, ctx->p.dims, check_mass, d.d_status, d.stream));
        ctx->launches++;
        const size_t shard_bytes = d.shard_count * 4 * ctx->esz, off = (size_t)d.rank * shard_bytes;
        if (ctx->ipc) {
            CU(cudaMemcpyAsync((char *)d.posm[1] + off, (char *)d.posm[0] + off, shard_bytes, cudaMemcpyDeviceToDevice, d.stream));
            for (int r = 0; r < ctx->world; ++r) {
                if (r == d.rank) continue;
                for (int k = 0; k < 2; ++k)
                    CU(cudaMemcpyAsync((char *)d.peer_posm[r][k] + off, (char *)d.posm[0] + off, shard_bytes, cudaMemcpyDefault, d.stream));
            }
        } else {
            NC(nccl().AllGather((char *)d.posm[0] + off, d.posm[0], shard_bytes, NCCL_UINT8, d.comm, d.stream));
            CU(cudaMemcpyAsync(d.posm[1], d.posm[0], full_bytes, cudaMemcpyDeviceToDevice, d.stream));
        }
    }
    if (ctx->ipc) {   // no peer may push into this rank's buffers before they are packed (and vice versa)
        int rc = ipc_barrier(ctx);
        if (rc != NBODY_OK) return rc;
    }
    for (Dev &d : ctx->devs) {
        CU(cudaSetDevice(d.device));
        CU(cudaStreamSynchronize(d.stream)); // `bodies` may be pageable and reused by the caller
        if (d.h_status && ((volatile unsigned *)d.h_status)[2]) {
            // the new state is on the device but must not be stepped with the uniform-mass kernel: report, and let a
            // later upload of equal masses (or a new context) make the context usable again
            ((volatile unsigned *)d.h_status)[2] = 0;
            set_err(ctx, "nbody_gpu_upload: masses changed; re-create the context (uniform-mass kernel in use)");
            return NBODY_ESTATE;
        }
    }
    return NBODY_OK;
}

ForceLaunch make_force(const nbody_ctx *c, const Dev &d, const Range &r, float dt)
{
    ForceLaunch L;
    memset(&L, 0, sizeof L);
    L.posm = d.posm[d.cur];
    L.dims = c->p.dims;
    L.small_tile = d.small_tile ? 1 : 0;
    L.uniform_mass = c->uniform ? 1 : 0;
    L.acc_scale = c->uniform ? c->p.G * c->uniform_mass : c->p.G;
    L.accp = d.accp;
    L.i_blk0 = (int)(d.shard_start / BLK);
    L.i_blk_local0 = 0;
    L.n_iblk = (int)(d.shard_count / BLK);
    L.n_iblk_shard = L.n_iblk;
    L.j_blk0 = r.j_blk0;
    L.j_nblk = r.j_nblk;
    L.j_body_limit = (long long)c->n;
    L.splits = r.splits;
    L.streamk_ctas = r.sk_ctas;
    L.slot0 = r.slot0;
    L.eps2 = c->p.eps * c->p.eps;                       // Quadtree ctor: e_sq = eps*eps in fp32
    L.eps2_f64 = (double)c->p.eps * (double)c->p.eps;
    L.fuse = 0;
    L.posm_next = d.posm[d.cur ^ 1];
    L.vel = d.vel;
    L.acc = d.acc;
    L.ip = make_ip(c, dt);
    return L;
}

// Enqueue one force evaluation (+ integration unless acc_only) on every local GPU.
int enqueue_step(nbody_ctx *ctx, float dt, bool acc_only, bool profile)
{
    const bool refc = !ctx->f64 && ctx->p.rsqrt_mode == NBODY_RSQRT_REFCOMPAT;
    // the kernel adds eps*eps (fp32) under flush-to-zero: guard against r^2 + eps^2 == 0 whenever that term can vanish
    const bool guard = (ctx->p.eps * ctx->p.eps < 1.17549435e-38f);
    const int par_prev = (int)((ctx->step_index + 1) & 1);     // parity of the previous step's push events
    // remote positions of the previous step must have landed: NCCL allgather done, or every peer's push done
    auto wait_remote = [&](Dev &d) -> cudaError_t {
        if (ctx->ipc) {
            ctx->launches++;
            return launch_wait_peer_flags(d.flags, ctx->world, d.rank, ctx->step_index, ctx->peer_timeout_ns,
                                          d.ipc_words + 1, d.stream);
        }
        if (!ctx->p2p) return cudaStreamWaitEvent(d.stream, d.ev_gathered, 0);
        for (Dev &o : ctx->devs) {
            if (&o == &d) continue;
            cudaError_t e = cudaStreamWaitEvent(d.stream, o.ev_pushed[par_prev], 0);
            if (e != cudaSuccess) return e;
        }
        return cudaSuccess;
    };
    for (Dev &d : ctx->devs) {
        CU(cudaSetDevice(d.device));
        const bool prof = profile && (&d == &ctx->devs[0]);
        if (prof) CU(cudaEventRecord(d.ev_t[0], d.stream));
        bool waited = false, bh_fused = false, col_grid_filled = false;
        if (ctx->bh) {
            if (d.gathered_pending) { CU(wait_remote(d)); waited = true; }
            int nl = 0;
            // small scenes on one GPU: the walk threads integrate their own targets (no separate integrator launch)
            bh_fused = ctx->world == 1 && !acc_only && !d.bh.warp_walk && ctx->n_padded <= 131072 && ctx->p.fuse_integrator != 0;
            static const bool fuse_insert_off = getenv("NBODY_BH_FUSE_INSERT") && atoi(getenv("NBODY_BH_FUSE_INSERT")) == 0;   // A/B switch
            // ... and fill the collision pass's screening grid (Simulation::step()); the grid is cleared HERE, ahead of the
            // build, so that nothing but kernels lies between the tree's last kernel and the walk (programmatic launches)
            const bool fill_grid = bh_fused && ctx->p.collide && !fuse_insert_off;
            if (fill_grid) CU(d.col.prepare(d.stream));
            CU(d.bh.build((const float *)d.posm[d.cur], ctx->n, d.stream, &nl));
            if (prof) {
                CU(cudaEventRecord(d.ev_t[3], d.stream));
                CU(cudaMemsetAsync(d.walk_visits, 0, 2 * sizeof(unsigned long long), d.stream));
            }
            BhFuseArgs fa;
            ColArgs col_a;
            ColGrid col_g;
            if (bh_fused) {
                fa.posm_next = (float *)d.posm[d.cur ^ 1]; fa.vel = (float *)d.vel; fa.acc = (float *)d.acc;
                fa.G = ctx->p.G; fa.ip = make_ip(ctx, dt);
                if (fill_grid) {
                    col_a = d.col.args((float *)d.posm[d.cur ^ 1], (float *)d.vel, ctx->n);
                    col_g = d.col.grid_view();
                    fa.col_args = &col_a; fa.col_grid = &col_g;
                    col_grid_filled = true;
                }
            }
            CU(d.bh.walk((const float *)d.posm[d.cur], ctx->n, ctx->p.theta, ctx->p.eps, refc, ctx->p.bh_fix_near_leaves != 0,
                         d.shard_start, d.shard_count, (float *)d.accp, prof ? d.walk_visits : nullptr, bh_fused ? &fa : nullptr, d.stream));
            ctx->launches += (unsigned long long)nl + 1;
        }
        for (const Range &r : d.plan) {
            if (ctx->bh) break;
            if (r.remote && d.gathered_pending && !waited) {
                CU(wait_remote(d));
                waited = true;
            }
            ForceLaunch L = make_force(ctx, d, r, dt);
            cudaError_t e;
            if (ctx->f64) e = launch_force_f64(L, d.stream);
            else if (refc) e = launch_force_f32_refcompat(L, d.stream);
            else {
                L.fuse = (d.fused && !acc_only) ? 1 : 0;
                e = launch_force_f32_fast(L, guard, d.stream);
            }
            CU(e);
            ctx->launches++;
        }
        if (d.gathered_pending && !waited) { // local-only plan cannot happen with world>1, but be safe
            CU(wait_remote(d));
        }
        if (prof) CU(cudaEventRecord(d.ev_t[1], d.stream));
        const bool fused_now = bh_fused || (d.fused && !acc_only && !ctx->f64 && !refc);
        if (!fused_now) {
            IntegLaunch I;
            memset(&I, 0, sizeof I);
            I.posm_cur = d.posm[d.cur];
            I.posm_next = d.posm[d.cur ^ 1];
            I.dests.n = 0;
            if (ctx->p2p && !acc_only) {   // integrate-and-push: new positions go to every GPU's next buffer
                for (Dev &o : ctx->devs) I.dests.p[I.dests.n++] = o.posm[o.cur ^ 1];
            } else if (ctx->ipc && !acc_only) {
                // the same across processes: every rank is at the same buffer parity (SPMD), and the kernel's last
                // CTA publishes "step_index + 1 steps pushed" into each peer's flag array
                I.dests.p[I.dests.n++] = d.posm[d.cur ^ 1];
                for (int r = 0; r < ctx->world; ++r) {
                    if (r == d.rank) continue;
                    I.dests.p[I.dests.n++] = d.peer_posm[r][d.cur ^ 1];
                    I.signal.slot[I.signal.n++] = d.peer_flags[r] + d.rank;
                }
                I.signal.value = ctx->step_index + 1;
                I.signal.arrive = d.ipc_words;
            } else {
                I.dests.p[I.dests.n++] = d.posm[d.cur ^ 1];
            }
            I.vel = d.vel;
            I.acc = d.acc;
            I.accp = d.accp;
            I.slots.n = 0;
            for (const Range &r : d.plan) {
                SlotRange &R = I.slots.r[I.slots.n++];
                R.slot0 = r.slot0; R.nslots = r.splits; R.sk_S = 0; R.sk_G = 0; R.sk_U = 0; R.tile_blks = 1;
                if (r.sk_ctas > 0) {
                    const int stage_blks = d.small_tile ? SMALL_STAGE_BLKS : FAST_STAGE_BLKS;
                    R.tile_blks = d.small_tile ? SMALL_TILE_BLKS : FAST_TILE_BLKS;
                    R.sk_S = (r.j_nblk + stage_blks - 1) / stage_blks;
                    R.sk_G = r.sk_ctas;
                    R.sk_U = (long long)((int)(d.shard_count / BLK) / R.tile_blks) * R.sk_S;
                }
            }
            I.i_blk0 = (int)(d.shard_start / BLK);
            I.n_iblk_shard = (int)(d.shard_count / BLK);
            I.acc_only = acc_only ? 1 : 0;
            I.n_real = (long long)ctx->n;
            const bool fastpath = !ctx->f64 && !refc && !ctx->bh;
            I.acc_scale = (fastpath && ctx->uniform) ? ctx->p.G * ctx->uniform_mass : ctx->p.G;
            I.ip = make_ip(ctx, dt);
            if (d.fused && acc_only) {
                // fused plans own no partial slot buffer semantics beyond slot 0: the non-fused
                // force launch above wrote slot 0.
                I.slots.n = 1;
                I.slots.r[0].nslots = 1;
            }
            CU(ctx->f64 ? launch_integrate_f64(I, d.stream) : launch_integrate_f32(I, d.stream));
            ctx->launches++;
        }
        if (prof) CU(cudaEventRecord(d.ev_t[4], d.stream));
        if (ctx->p.collide && !acc_only) { // Simulation::step(): iterate(dt) ; collide()
            int nl = 0;
            CU(d.col.run((float *)d.posm[d.cur ^ 1], (float *)d.vel, ctx->n, d.stream, &nl, col_grid_filled));
            ctx->launches += (unsigned long long)nl;
        }
        if (prof) CU(cudaEventRecord(d.ev_t[2], d.stream));
        if (!acc_only) {
            size_t real = 0;
            if (d.shard_start < ctx->n) real = std::min(ctx->n - d.shard_start, d.shard_count);
            ctx->interactions += (unsigned long long)real * (unsigned long long)ctx->n;
        }
    }
    if (acc_only) return NBODY_OK;

    if (ctx->ipc) {
        for (Dev &d : ctx->devs) d.gathered_pending = true;   // pushed by the integrator kernel; peers' flags gate the next step
    } else if (ctx->p2p) {
        // the integrator kernels already stored the new positions into every peer's next buffer
        const int par = (int)(ctx->step_index & 1);
        for (Dev &d : ctx->devs) {
            CU(cudaSetDevice(d.device));
            CU(cudaEventRecord(d.ev_pushed[par], d.stream));
            d.gathered_pending = true;
        }
    } else if (ctx->world > 1) {
        // new positions of every shard -> every GPU, in place in the next buffer, on the comm stream
        for (Dev &d : ctx->devs) {
            CU(cudaSetDevice(d.device));
            CU(cudaEventRecord(d.ev_integrated, d.stream));
            CU(cudaStreamWaitEvent(d.comm_stream, d.ev_integrated, 0));
        }
        NC(nccl().GroupStart());
        for (Dev &d : ctx->devs) {
            const size_t bytes = d.shard_count * 4 * ctx->esz;
            char *base = (char *)d.posm[d.cur ^ 1];
            int r = nccl().AllGather(base + (size_t)d.rank * bytes, base, bytes, NCCL_UINT8, d.comm, d.comm_stream);
            if (r != 0) { nccl().GroupEnd(); NC(r); }
        }
        NC(nccl().GroupEnd());
        for (Dev &d : ctx->devs) {
            CU(cudaSetDevice(d.device));
            CU(cudaEventRecord(d.ev_gathered, d.comm_stream));
            d.gathered_pending = true;
        }
    }
    for (Dev &d : ctx->devs) d.cur ^= 1;
    ctx->step_index++;
    return NBODY_OK;
}

int sync_all(nbody_ctx *ctx)
{
    for (Dev &d : ctx->devs) {
        CU(cudaSetDevice(d.device));
        if (ctx->ipc && d.gathered_pending) {   // this rank's buffers are complete only once every peer's push has landed
            CU(launch_wait_peer_flags(d.flags, ctx->world, d.rank, ctx->step_index, ctx->peer_timeout_ns,
                                      d.ipc_words + 1, d.stream));
            ctx->launches++;
        }
        if (d.comm_stream) CU(cudaStreamSynchronize(d.comm_stream));
        CU(cudaStreamSynchronize(d.stream));
        d.gathered_pending = false;
        if (d.h_status) {
            const volatile unsigned *hs = d.h_status;
            if (hs[0]) {
                set_err(ctx, "Barnes-Hut tree needed more cells than reserved (%u): forces were truncated. Clustered bodies form long "
                             "single-child chains; raise NBODY_BH_NODE_FACTOR (cells reserved per body, default 4, at most 33)", d.bh.node_cap);
                return NBODY_ENOMEM;
            }
            if (hs[1]) {
                set_err(ctx, "collision pass: cell-entry or pair buffer overflow (a body spans more than 4096 grid cells or 65536 cell strips, or a dense "
                             "clump produced more than %u pairs): the pass was abandoned for at least one step", d.col.pair_cap);
                return NBODY_ESTATE;
            }
        }
        if (ctx->ipc) {
            unsigned status = 0;
            CU(cudaMemcpy(&status, d.ipc_words + 1, sizeof status, cudaMemcpyDeviceToHost));
            if (status) {
                set_err(ctx, "peer exchange: a peer rank did not publish its positions within %.0f s", ctx->peer_timeout_ns * 1e-9);
                return NBODY_ESTATE;
            }
        }
    }
    return NBODY_OK;
}

// One process per GPU on one node: map every peer's two position buffers and flag array into this process
// (cudaIpcGetMemHandle / cudaIpcOpenMemHandle, handles exchanged through the NCCL communicator), so that the
// integrator kernel can store new positions straight into the peers' HBM over NVLink.  All ranks agree on
// the outcome (sum of per-rank success flags); on failure everything is unmapped and the ncclAllGather path
// stays in use, unless exchange == 2 demanded the peer path.
int setup_ipc(nbody_ctx *ctx)
{
    struct Blob { cudaIpcMemHandle_t h[3]; int ok; int pad[15]; };
    static_assert(sizeof(Blob) == 256, "blob layout");
    Dev &d = ctx->devs[0];
    const int W = ctx->world, me = d.rank;
    CU(cudaSetDevice(d.device));
    CU(cudaMalloc(&d.flags, NBODY_MAX_GPUS * sizeof(unsigned long long)));
    CU(cudaMemset(d.flags, 0, NBODY_MAX_GPUS * sizeof(unsigned long long)));
    CU(cudaMalloc(&d.ipc_words, 64));
    CU(cudaMemset(d.ipc_words, 0, 64));
    Blob mine;
    memset(&mine, 0, sizeof mine);
    mine.ok = 1;
    void *exported[3] = {d.posm[0], d.posm[1], d.flags};
    for (int k = 0; k < 3; ++k)
        if (cudaIpcGetMemHandle(&mine.h[k], exported[k]) != cudaSuccess) { cudaGetLastError(); mine.ok = 0; }
    Blob *dev_blobs = nullptr;
    CU(cudaMalloc(&dev_blobs, (size_t)W * sizeof(Blob)));
    std::vector<Blob> blobs(W);
    cudaError_t e = cudaMemcpyAsync(dev_blobs + me, &mine, sizeof mine, cudaMemcpyHostToDevice, d.stream);
    int nr = 0;
    if (e == cudaSuccess) nr = nccl().AllGather(dev_blobs + me, dev_blobs, sizeof(Blob), NCCL_UINT8, d.comm, d.stream);
    if (e == cudaSuccess && nr == 0) e = cudaMemcpyAsync(blobs.data(), dev_blobs, (size_t)W * sizeof(Blob), cudaMemcpyDeviceToHost, d.stream);
    if (e == cudaSuccess && nr == 0) e = cudaStreamSynchronize(d.stream);
    if (nr != 0) { cudaFree(dev_blobs); NC(nr); }
    if (e != cudaSuccess) { cudaFree(dev_blobs); CU(e); }

    int ok = 1;
    for (int r = 0; r < W; ++r) ok &= blobs[r].ok;
    for (int r = 0; r < W && ok; ++r) {
        if (r == me) continue;
        void *ptr[3] = {nullptr, nullptr, nullptr};
        for (int k = 0; k < 3 && ok; ++k)
            if (cudaIpcOpenMemHandle(&ptr[k], blobs[r].h[k], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                set_err(ctx, "cudaIpcOpenMemHandle(rank %d): %s", r, cudaGetErrorString(cudaGetLastError()));
                ok = 0;
            }
        d.peer_posm[r][0] = ptr[0];
        d.peer_posm[r][1] = ptr[1];
        d.peer_flags[r] = (unsigned long long *)ptr[2];
    }
    // consensus: every rank must have mapped every peer
    int *dev_ok = reinterpret_cast<int *>(dev_blobs);
    int sum = 0;
    e = cudaMemcpyAsync(dev_ok, &ok, sizeof ok, cudaMemcpyHostToDevice, d.stream);
    if (e == cudaSuccess) nr = nccl().AllReduce(dev_ok, dev_ok, 1, NCCL_INT32, NCCL_SUM, d.comm, d.stream);
    if (e == cudaSuccess && nr == 0) e = cudaMemcpyAsync(&sum, dev_ok, sizeof sum, cudaMemcpyDeviceToHost, d.stream);
    if (e == cudaSuccess && nr == 0) e = cudaStreamSynchronize(d.stream);
    cudaFree(dev_blobs);
    if (nr != 0) NC(nr);
    CU(e);
    ctx->ipc = (sum == W);
    if (!ctx->ipc) {
        for (int r = 0; r < W; ++r) {
            for (int k = 0; k < 2; ++k) if (d.peer_posm[r][k]) { cudaIpcCloseMemHandle(d.peer_posm[r][k]); d.peer_posm[r][k] = nullptr; }
            if (d.peer_flags[r]) { cudaIpcCloseMemHandle(d.peer_flags[r]); d.peer_flags[r] = nullptr; }
        }
        cudaGetLastError();
        if (ctx->p.exchange == 2) {
            if (!ctx->err[0]) set_err(ctx, "peer exchange requested (exchange = 2) but a rank could not map its peers");
            return NBODY_ECUDA;
        }
        ctx->err[0] = 0;
    }
    if (const char *t = getenv("NBODY_PEER_TIMEOUT_S")) {
        const double sec = atof(t);
        if (sec > 0) ctx->peer_timeout_ns = (unsigned long long)(sec * 1e9);
    }
    return NBODY_OK;
}

// Cross-rank barrier in stream order (a one-word allreduce): no kernel enqueued after it on any rank starts
// before every rank has reached it.  Separates "all ranks have (re)written their buffers" from the first push.
int ipc_barrier(nbody_ctx *ctx)
{
    Dev &d = ctx->devs[0];
    CU(cudaSetDevice(d.device));
    int *w = reinterpret_cast<int *>(d.ipc_words + 4);
    NC(nccl().AllReduce(w, w, 1, NCCL_INT32, NCCL_SUM, d.comm, d.stream));
    return NBODY_OK;
}

void free_all(nbody_ctx *c)
{
    for (Dev &d : c->devs) {
        cudaSetDevice(d.device);
        for (int r = 0; r < NBODY_MAX_GPUS; ++r) {
            for (int k = 0; k < 2; ++k) if (d.peer_posm[r][k]) cudaIpcCloseMemHandle(d.peer_posm[r][k]);
            if (d.peer_flags[r]) cudaIpcCloseMemHandle(d.peer_flags[r]);
        }
        if (d.flags) cudaFree(d.flags);
        if (d.ipc_words) cudaFree(d.ipc_words);
        if (d.comm && nccl().CommDestroy) nccl().CommDestroy(d.comm);
        for (int k = 0; k < 2; ++k) if (d.posm[k]) cudaFree(d.posm[k]);
        if (d.vel) cudaFree(d.vel);
        if (d.acc) cudaFree(d.acc);
        if (d.accp) cudaFree(d.accp);
        if (d.aos) cudaFree(d.aos);
        if (d.energy5) cudaFree(d.energy5);
        if (d.walk_visits) cudaFree(d.walk_visits);
        if (d.h_status) cudaFreeHost(d.h_status);
        d.bh.release();
        d.col.release();
        if (d.graph) cudaGraphExecDestroy(d.graph);
        for (int k = 0; k < 2; ++k) if (d.graph1[k]) cudaGraphExecDestroy(d.graph1[k]);
        if (d.ev_integrated) cudaEventDestroy(d.ev_integrated);
        if (d.ev_gathered) cudaEventDestroy(d.ev_gathered);
        for (int k = 0; k < 2; ++k) if (d.ev_pushed[k]) cudaEventDestroy(d.ev_pushed[k]);
        for (int k = 0; k < 6; ++k) if (d.ev_t[k]) cudaEventDestroy(d.ev_t[k]);
        if (d.comm_stream) cudaStreamDestroy(d.comm_stream);
        if (d.own_stream && d.stream) cudaStreamDestroy(d.stream);
    }
    if (c->h_stage) cudaFreeHost(c->h_stage);
    delete c;
}

} // namespace

// ================================================================================================
extern "C" {

void nbody_params_default(nbody_params *p)
{
    if (!p) return;
    memset(p, 0, sizeof *p);
    p->struct_size = (uint32_t)sizeof *p;
    p->dims = 2;                    // the reference is 2-D (Vec2.hpp:17-20)
    p->eps = 1.0f;                  // Simulation.hpp:59
    p->G = 1.0f;
    p->precision = NBODY_PRECISION_F32;
    p->rsqrt_mode = NBODY_RSQRT_FAST;
    p->force_algo = NBODY_FORCE_ALLPAIRS;
    p->theta = 1.0f;                // Simulation.hpp:59
    p->integ_flags = 0;
    p->max_velocity = 1000.0f;      // Simulation.hpp:124
    p->boundary_radius = 100000.0f; // :120
    p->soft_boundary = 0.8f;        // :121
    p->boundary_force = 0.9f;       // :122
    p->damping = 0.9995f;           // :123
    p->j_splits = 0;
    p->fuse_integrator = -1;
    p->use_graph = -1;
    p->force_variant = -1;
    p->ngpus = 1;
    p->world = 1;
    p->rank = 0;
    p->stream = nullptr;
}

int nbody_gpu_init(nbody_ctx **out, const nbody_params *p, const nbody_body_t *bodies, size_t n)
{
    nbody_ctx *ctx = nullptr; // for the CU/NC macros before the context exists
    if (!out) return NBODY_EINVAL;
    *out = nullptr;
    if (!p || !bodies || n == 0 || p->struct_size != sizeof(nbody_params)) {
        set_err(nullptr, "nbody_gpu_init: null argument, n == 0 or nbody_params size mismatch");
        return NBODY_EINVAL;
    }
    if ((p->dims != 2 && p->dims != 3) || !(p->eps >= 0.0f) ||
        (p->precision != NBODY_PRECISION_F32 && p->precision != NBODY_PRECISION_F64) ||
        (p->rsqrt_mode != NBODY_RSQRT_FAST && p->rsqrt_mode != NBODY_RSQRT_REFCOMPAT) ||
        p->ngpus > NBODY_MAX_GPUS || n > ((size_t)1 << 30)) {
        set_err(nullptr, "nbody_gpu_init: invalid parameter");
        return NBODY_EINVAL;
    }
    if (p->exchange < 0 || p->exchange > 2 || p->sort_impl < 0 || p->sort_impl > 2 || p->bh_walk < 0 || p->bh_walk > 2 ||
        !(p->theta >= 0.0f) || p->world < 0 || p->ngpus < 0) {
        set_err(nullptr, "nbody_gpu_init: exchange, sort_impl and bh_walk take 0..2; theta, world and ngpus must not be negative");
        return NBODY_EINVAL;
    }
    if (p->force_algo != NBODY_FORCE_ALLPAIRS && p->force_algo != NBODY_FORCE_BARNES_HUT) {
        set_err(nullptr, "nbody_gpu_init: unknown force_algo %d", p->force_algo);
        return NBODY_EINVAL;
    }
    if (p->collide && (p->dims != 2 || p->precision != NBODY_PRECISION_F32 || std::max(1, p->ngpus) > 1 || p->world > 1)) {
        set_err(nullptr, "nbody_gpu_init: the collision pass is the reference's 2-D fp32 collide() on one GPU");
        return NBODY_EINVAL;
    }
    if (p->force_algo == NBODY_FORCE_BARNES_HUT && p->precision != NBODY_PRECISION_F32) {
        set_err(nullptr, "nbody_gpu_init: the Barnes-Hut path is fp32 (the reference's quadtree in 2-D, its octree generalisation in 3-D)");
        return NBODY_EINVAL;
    }
    const int nlocal = std::max(1, p->ngpus);
    const bool multiproc = p->world > 1;
    if (multiproc && nlocal > 1) {
        set_err(nullptr, "nbody_gpu_init: use either ngpus>1 (one process) or world>1 (one process per GPU)");
        return NBODY_EINVAL;
    }
    const int world = multiproc ? p->world : nlocal;
    if (multiproc && (p->rank < 0 || p->rank >= p->world)) return NBODY_EINVAL;
    if (multiproc && p->world > NBODY_MAX_GPUS && p->exchange != 1) {
        // the peer exchange keeps per-rank tables of NBODY_MAX_GPUS entries (mapped buffers, flag words, kernel arguments)
        set_err(nullptr, "nbody_gpu_init: world = %d exceeds NBODY_MAX_GPUS = %d; larger worlds need exchange = 1 (ncclAllGather)",
                p->world, NBODY_MAX_GPUS);
        return NBODY_EINVAL;
    }

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        set_err(nullptr, "nbody_gpu_init: no CUDA device visible");
        return NBODY_ENODEV;
    }
    ctx = new (std::nothrow) nbody_ctx();
    if (!ctx) return NBODY_ENOMEM;
    ctx->p = *p;
    ctx->n = n;
    ctx->world = world;
    ctx->f64 = (p->precision == NBODY_PRECISION_F64);
    ctx->bh = (p->force_algo == NBODY_FORCE_BARNES_HUT);
    ctx->esz = ctx->f64 ? 8 : 4;
    if (const char *ff = getenv("NBODY_FORCE_FORM")) ctx->streamk = strcmp(ff, "split") != 0;
    {
        // Geometry and padding.  The fast kernel's 2048-target tiles need n padded to 2048 per rank; a small
        // shard, for which those tiles could not give every SM two CTAs, runs 512-target tiles and pads to 512
        // (at the reference's own n = 25,000 that is 25,088 bodies instead of 26,624: 12 % less work).
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, p->device_ids[0]);
        const size_t per_l = (size_t)TARGET_GRANULE * (size_t)world;
        const size_t np_l = ((n + per_l - 1) / per_l) * per_l;
        const bool fastpath = !ctx->f64 && !ctx->bh && p->rsqrt_mode == NBODY_RSQRT_FAST;
        const long long units_l = (long long)(np_l / world / TARGET_GRANULE) * std::max<long long>(1, (long long)(np_l / BLK) / 8);
        ctx->small_tile = fastpath && units_l < 2LL * sms && !(p->fuse_integrator == 1 && p->j_splits == 1);
        // refcompat / fp64 / Barnes-Hut work per 128- or 256-target tile: pad to one block per rank
        const size_t per_s = (size_t)(fastpath ? SMALL_TILE_BLKS : 1) * BLK * (size_t)world;
        ctx->n_padded = (ctx->small_tile || !fastpath) ? ((n + per_s - 1) / per_s) * per_s : np_l;
    }
    if (!ctx->f64 && !ctx->bh && p->rsqrt_mode == NBODY_RSQRT_FAST && p->force_variant != 0) {
        // uniform-mass form: valid when every body has the same positive mass (bit-equal), so that
        // sum_j m_j f(r_ij) == m * sum_j f(r_ij) term for term.  force_variant = 0 disables it.
        const float m0 = bodies[0].mass;
        bool same = m0 > 0.0f;
        for (size_t i = 1; i < n && same; ++i) same = (bodies[i].mass == m0);
        ctx->uniform = same;
        ctx->uniform_mass = same ? m0 : 0.0f;
    }
    int rc = NBODY_OK;
    auto fail = [&](int code) {
        memcpy(g_init_err, ctx->err, sizeof g_init_err);
        free_all(ctx);
        return code;
    };

    ctx->devs.resize(nlocal);
    for (int k = 0; k < nlocal; ++k) {
        Dev &d = ctx->devs[k];
        d.device = p->device_ids[k];
        if (d.device < 0 || d.device >= ndev) {
            set_err(ctx, "nbody_gpu_init: device ordinal %d out of range (0..%d)", d.device, ndev - 1);
            return fail(NBODY_ENODEV);
        }
        d.rank = multiproc ? p->rank : k;
        d.shard_count = ctx->n_padded / (size_t)world;
        d.shard_start = d.shard_count * (size_t)d.rank;
        if (p->stream && nlocal == 1) { d.stream = (cudaStream_t)p->stream; d.own_stream = false; }
    }
    {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, ctx->devs[0].device) != cudaSuccess) {
            set_err(ctx, "cudaGetDeviceProperties failed");
            return fail(NBODY_ECUDA);
        }
        if (prop.major < 10) {
            set_err(ctx, "device %d is sm_%d%d; this library contains sm_100a code only", ctx->devs[0].device,
                    prop.major, prop.minor);
            return fail(NBODY_ENODEV);
        }
        ctx->sm_count = prop.multiProcessorCount;
        int khz = 0;
        cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, ctx->devs[0].device);
        ctx->sm_clock_khz = khz;
        cudaSetDevice(ctx->devs[0].device);
        ctx->ctas_per_sm = force_f32_fast_ctas_per_sm(ctx->uniform, false);
        ctx->ctas_per_sm_small = force_f32_fast_ctas_per_sm(ctx->uniform, true);
        ctx->sk_ctas_per_sm = force_f32_streamk_ctas_per_sm(ctx->uniform, false);
        ctx->sk_ctas_per_sm_small = force_f32_streamk_ctas_per_sm(ctx->uniform, true);
    }
    if (!multiproc && nlocal > 1 && p->exchange != 1) {
        // peer-to-peer exchange needs every local GPU to reach every other one (NVLink / NVSwitch)
        bool all = true;
        for (int a = 0; a < nlocal && all; ++a)
            for (int b = 0; b < nlocal && all; ++b) {
                if (a == b) continue;
                int can = 0;
                if (cudaDeviceCanAccessPeer(&can, ctx->devs[a].device, ctx->devs[b].device) != cudaSuccess || !can) all = false;
            }
        if (all) {
            for (int a = 0; a < nlocal; ++a) {
                cudaSetDevice(ctx->devs[a].device);
                for (int b = 0; b < nlocal; ++b) {
                    if (a == b) continue;
                    cudaError_t pe = cudaDeviceEnablePeerAccess(ctx->devs[b].device, 0);
                    if (pe == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
                    else if (pe != cudaSuccess) all = false;
                }
            }
        }
        ctx->p2p = all;
    }
    for (Dev &d : ctx->devs) {
        if ((rc = plan_device(ctx, d)) != NBODY_OK) return fail(rc);
        if ((rc = alloc_device(ctx, d)) != NBODY_OK) return fail(rc);
    }
    if (cudaMallocHost(&ctx->h_stage, ctx->n * sizeof(nbody_body_t)) != cudaSuccess) {
        set_err(ctx, "cudaMallocHost(%zu) failed", ctx->n * sizeof(nbody_body_t));
        return fail(NBODY_ENOMEM);
    }
    if (world > 1 && !ctx->p2p) {
        if (!nccl().load()) {
            set_err(ctx, "libnccl.so.2 could not be loaded: %s", dlerror());
            return fail(NBODY_ENCCL);
        }
        if (multiproc) {
            ncclUniqueId id;
            memcpy(&id, p->nccl_id, sizeof id);
            cudaSetDevice(ctx->devs[0].device);
            int r = nccl().CommInitRank(&ctx->devs[0].comm, world, id, p->rank);
            if (r != 0) { set_err(ctx, "ncclCommInitRank: %s", nccl().GetErrorString(r)); return fail(NBODY_ENCCL); }
        } else {
            std::vector<ncclComm_t> comms(nlocal);
            std::vector<int> ids(nlocal);
            for (int k = 0; k < nlocal; ++k) ids[k] = ctx->devs[k].device;
            int r = nccl().CommInitAll(comms.data(), nlocal, ids.data());
            if (r != 0) { set_err(ctx, "ncclCommInitAll: %s", nccl().GetErrorString(r)); return fail(NBODY_ENCCL); }
            for (int k = 0; k < nlocal; ++k) ctx->devs[k].comm = comms[k];
        }
    }
    if (multiproc && p->exchange != 1) {
        if ((rc = setup_ipc(ctx)) != NBODY_OK) return fail(rc);
    }
    if ((rc = upload_state(ctx, bodies, false, false)) != NBODY_OK) return fail(rc);
    *out = ctx;
    return NBODY_OK;
}

// Launch-bound regime (small N: a step is two short kernels): replay pairs of steps from a CUDA graph.
static int step_with_graph(nbody_ctx *ctx, float dt, int &remaining)
{
    Dev &d = ctx->devs[0];
    CU(cudaSetDevice(d.device));
    if (!d.graph || d.graph_dt != dt || d.graph_cur != d.cur) {
        if (d.graph) { cudaGraphExecDestroy(d.graph); d.graph = nullptr; }
        const unsigned long long l0 = ctx->launches, i0 = ctx->interactions;
        cudaGraph_t g = nullptr;
        CU(cudaStreamBeginCapture(d.stream, cudaStreamCaptureModeThreadLocal));
        int rc = enqueue_step(ctx, dt, false, false);
        if (rc == NBODY_OK) rc = enqueue_step(ctx, dt, false, false);
        cudaError_t e = cudaStreamEndCapture(d.stream, &g);
        d.graph_launches = ctx->launches - l0;            // nothing ran yet: undo the bookkeeping of the capture
        d.graph_interactions = ctx->interactions - i0;
        ctx->launches = l0;
        ctx->interactions = i0;
        if (rc != NBODY_OK) { if (g) cudaGraphDestroy(g); return rc; }
        CU(e);
        e = cudaGraphInstantiate(&d.graph, g, 0);
        cudaGraphDestroy(g);
        CU(e);
        d.graph_dt = dt;
        d.graph_cur = d.cur;                              // two steps flip the buffer parity twice
    }
    while (remaining >= 2) {
        CU(cudaGraphLaunch(d.graph, d.stream));
        ctx->launches += d.graph_launches;
        ctx->interactions += d.graph_interactions;
        remaining -= 2;
    }
    return NBODY_OK;
}

// The Barnes-Hut cell reservation follows the tree: every build reports its cell count through the mapped status words,
// and before new steps are enqueued a reservation that is more than 3/4 full is doubled (up to the 33 cells per body a
// tree can need), so that a scene whose bodies cluster over time does not run into the overflow error.
static int grow_bh_reservation_if_needed(nbody_ctx *ctx)
{
    for (Dev &d : ctx->devs) {
        if (!d.h_status) continue;
        const unsigned cells = ((volatile unsigned *)d.h_status)[4];
        const unsigned long long limit = 33ull * ctx->n + 1024ull;
        if ((unsigned long long)cells * 4ull <= (unsigned long long)d.bh.node_cap * 3ull || d.bh.node_cap >= limit) continue;
        CU(cudaSetDevice(d.device));
        CU(cudaStreamSynchronize(d.stream));
        const double factor = std::min(33.0, 2.0 * (double)d.bh.node_cap / (double)ctx->n);
        const int cluster_mode = d.bh.cluster_ctas > 0 ? 2 : 1;
        const bool warp_walk = d.bh.warp_walk;
        const unsigned window = d.bh.walk_window;
        unsigned *status = d.bh.status;
        d.bh.release();
        CU(d.bh.alloc(ctx->n, ctx->p.dims, factor, cluster_mode, (size_t)-1));
        d.bh.status = status; d.bh.warp_walk = warp_walk; d.bh.walk_window = window;
        if (d.graph) { cudaGraphExecDestroy(d.graph); d.graph = nullptr; }       // the graphs hold the old buffers
        for (int k = 0; k < 2; ++k) if (d.graph1[k]) { cudaGraphExecDestroy(d.graph1[k]); d.graph1[k] = nullptr; }
        ((volatile unsigned *)d.h_status)[4] = 0;
    }
    return NBODY_OK;
}

// One step per call (nbody_gpu_step(ctx, dt, 1) from a viewer loop or an upload / step / download cycle): one graph per
// buffer parity, so such a caller pays one graph launch instead of ~18 kernel launches per step.
static int step_with_single_graph(nbody_ctx *ctx, float dt)
{
    Dev &d = ctx->devs[0];
    CU(cudaSetDevice(d.device));
    const int par = d.cur;
    if (!d.graph1[par] || d.graph1_dt[par] != dt) {
        if (d.graph1[par]) { cudaGraphExecDestroy(d.graph1[par]); d.graph1[par] = nullptr; }
        const unsigned long long l0 = ctx->launches, i0 = ctx->interactions, s0 = ctx->step_index;
        cudaGraph_t g = nullptr;
        CU(cudaStreamBeginCapture(d.stream, cudaStreamCaptureModeThreadLocal));
        const int rc = enqueue_step(ctx, dt, false, false);
        cudaError_t e = cudaStreamEndCapture(d.stream, &g);
        d.graph1_launches = ctx->launches - l0;           // nothing ran yet: undo the bookkeeping of the capture
        d.graph1_interactions = ctx->interactions - i0;
        ctx->launches = l0;
        ctx->interactions = i0;
        ctx->step_index = s0;
        d.cur = par;
        if (rc != NBODY_OK) { if (g) cudaGraphDestroy(g); return rc; }
        CU(e);
        e = cudaGraphInstantiate(&d.graph1[par], g, 0);
        cudaGraphDestroy(g);
        CU(e);
        d.graph1_dt[par] = dt;
    }
    CU(cudaGraphLaunch(d.graph1[par], d.stream));
    ctx->launches += d.graph1_launches;
    ctx->interactions += d.graph1_interactions;
    d.cur ^= 1;
    ctx->step_index++;
    return NBODY_OK;
}

int nbody_gpu_step(nbody_ctx *ctx, float dt, int nsteps)
{
    if (!ctx || nsteps < 0 || !(dt == dt)) return NBODY_EINVAL;
    // auto: graphs pay off only when the step is launch-bound (small shards) and the call is long enough
    // (the Barnes-Hut build and the collision pass are fully asynchronous -- sorts, scan, COM pass, pair
    //  discovery and resolve keep their counters on the device -- so they capture too)
    if (ctx->bh) {
        const int rc = grow_bh_reservation_if_needed(ctx);
        if (rc != NBODY_OK) return rc;
    }
    const bool graph_ok = ctx->world == 1 && ctx->devs.size() == 1 && !ctx->profile_next;
    const bool want = ctx->p.use_graph == 1 || (ctx->p.use_graph < 0 && (ctx->n_padded <= 32768 || (ctx->bh && ctx->n_padded <= 262144)));
    if (graph_ok && want && nsteps >= 8) {
        int rc = step_with_graph(ctx, dt, nsteps);
        if (rc != NBODY_OK) return rc;
    }
    // short calls: once a context has been stepped a few times in calls of fewer than 8 steps (a per-frame caller), its
    // steps replay from the one-step graphs as well
    if (nsteps > 0 && nsteps < 8) ctx->short_calls++;
    if (graph_ok && want && nsteps > 0 && nsteps < 8 && (ctx->short_calls > 3 || ctx->p.use_graph == 1)) {
        for (; nsteps > 0; --nsteps) {
            int rc = step_with_single_graph(ctx, dt);
            if (rc != NBODY_OK) return rc;
        }
    }
    for (int s = 0; s < nsteps; ++s) {
        const bool prof = ctx->profile_next && s == 0;
        int rc = enqueue_step(ctx, dt, false, prof);
        if (rc != NBODY_OK) return rc;
        if (prof) {
            Dev &d = ctx->devs[0];
            CU(cudaSetDevice(d.device));
            CU(cudaEventSynchronize(d.ev_t[2]));
            CU(cudaEventElapsedTime(&ctx->last_force_ms, d.ev_t[0], d.ev_t[1]));
            CU(cudaEventElapsedTime(&ctx->last_integ_ms, d.ev_t[1], d.ev_t[4]));
            CU(cudaEventElapsedTime(&ctx->last_collide_ms, d.ev_t[4], d.ev_t[2]));
            ctx->last_build_ms = 0.f;
            ctx->last_visits = 0; ctx->last_visits_max = 0;
            if (ctx->bh) {
                CU(cudaEventElapsedTime(&ctx->last_build_ms, d.ev_t[0], d.ev_t[3]));
                CU(cudaMemcpy(&ctx->last_visits, d.walk_visits, sizeof(unsigned long long), cudaMemcpyDeviceToHost));
                CU(cudaMemcpy(&ctx->last_visits_max, d.walk_visits + 1, sizeof(unsigned long long), cudaMemcpyDeviceToHost));
            }
            ctx->profile_next = 0;
        }
    }
    return NBODY_OK;
}

int nbody_gpu_accel_only(nbody_ctx *ctx)
{
    if (!ctx) return NBODY_EINVAL;
    return enqueue_step(ctx, 0.0f, true, false);
}

int nbody_gpu_sync(nbody_ctx *ctx)
{
    if (!ctx) return NBODY_EINVAL;
    return sync_all(ctx);
}

int nbody_gpu_download(nbody_ctx *ctx, nbody_body_t *bodies, size_t n, unsigned fields)
{
    if (!ctx || !bodies || n != ctx->n || (fields & ~NBODY_FIELD_ALL) || fields == 0) return NBODY_EINVAL;
    int rc = sync_all(ctx);
    if (rc != NBODY_OK) return rc;
    for (Dev &d : ctx->devs) {
        if (d.shard_start >= ctx->n) continue;
        const size_t cnt = std::min(ctx->n - d.shard_start, d.shard_count);
        CU(cudaSetDevice(d.device));
        CU(launch_unpack(d.aos, ctx->n, d.shard_start, d.shard_count, d.posm[d.cur], d.vel, d.acc, ctx->f64, d.stream));
        ctx->launches++;
        nbody_body_t *dst = (fields == NBODY_FIELD_ALL) ? bodies + d.shard_start : ctx->h_stage + d.shard_start;
        CU(cudaMemcpyAsync(dst, d.aos, cnt * sizeof(nbody_body_t), cudaMemcpyDeviceToHost, d.stream));
    }
    for (Dev &d : ctx->devs) {
        CU(cudaSetDevice(d.device));
        CU(cudaStreamSynchronize(d.stream));
    }
    if (fields != NBODY_FIELD_ALL) {
        // merge the requested fields from the pinned staging copy; the caller's other fields stay untouched.
        // Memory-bound host work: split over a few threads when the shard is large (a viewer polling positions).
        const bool z = ctx->p.dims == 3;
        const nbody_body_t *stage = ctx->h_stage;
        auto merge = [=](size_t i0, size_t i1) {
            for (size_t i = i0; i < i1; ++i) {
                const nbody_body_t &s = stage[i];
                nbody_body_t &t = bodies[i];
                if (fields & NBODY_FIELD_POS) { t.pos[0] = s.pos[0]; t.pos[1] = s.pos[1]; if (z) t.pos_z = s.pos_z; }
                if (fields & NBODY_FIELD_VEL) { t.vel[0] = s.vel[0]; t.vel[1] = s.vel[1]; if (z) t.vel_z = s.vel_z; }
                if (fields & NBODY_FIELD_ACC) { t.acc[0] = s.acc[0]; t.acc[1] = s.acc[1]; if (z) t.acc_z = s.acc_z; }
            }
        };
        for (Dev &d : ctx->devs) {
            if (d.shard_start >= ctx->n) continue;
            const size_t i0 = d.shard_start, i1 = std::min(ctx->n, d.shard_start + d.shard_count), cnt = i1 - i0;
            const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
            const size_t nt = std::min<size_t>(std::min<size_t>(hw, 8), cnt / 131072);
            if (nt <= 1) { merge(i0, i1); continue; }
            std::vector<std::thread> pool;
            try {
                for (size_t t = 1; t < nt; ++t) pool.emplace_back(merge, i0 + cnt * t / nt, i0 + cnt * (t + 1) / nt);
            } catch (...) {               // could not start a thread: this one does the rest
                const size_t done = pool.size() + 1;
                merge(i0 + cnt * done / nt, i1);
            }
            merge(i0, i0 + cnt / nt);
            for (std::thread &th : pool) th.join();
        }
    }
    return NBODY_OK;
}

int nbody_gpu_upload(nbody_ctx *ctx, const nbody_body_t *bodies, size_t n)
{
    if (!ctx || !bodies || n != ctx->n) return NBODY_EINVAL;
    int rc = sync_all(ctx);
    if (rc != NBODY_OK) return rc;
    // one process per GPU: only this rank's shard is read and crosses PCIe; the exchange carries it to the peers.
    // (the uniform-mass kernel stays valid only while the masses stay as uploaded: checked by the pack kernel, on the
    //  device, instead of a host loop over every record)
    return upload_state(ctx, bodies, ctx->p.world > 1, true);
}

int nbody_gpu_download_f64(nbody_ctx *ctx, double *pos3, double *vel3, double *acc3, size_t n)
{
    if (!ctx || n != ctx->n) return NBODY_EINVAL;
    int rc = sync_all(ctx);
    if (rc != NBODY_OK) return rc;
    for (Dev &d : ctx->devs) {
        if (d.shard_start >= ctx->n) continue;
        const size_t cnt = std::min(ctx->n - d.shard_start, d.shard_count);
        CU(cudaSetDevice(d.device));
        // reuse the AoS staging area (64 B/body >= 3 x 24 B/body? no: 72 B) -> separate scratch
        double *scratch = nullptr;
        CU(cudaMalloc(&scratch, cnt * 9 * sizeof(double)));
        double *dp = scratch, *dv = scratch + 3 * cnt, *da = scratch + 6 * cnt;
        cudaError_t e = launch_unpack_f64(pos3 ? dp : nullptr, vel3 ? dv : nullptr, acc3 ? da : nullptr, ctx->n,
                                          d.shard_start, d.shard_count, d.posm[d.cur], d.vel, d.acc, ctx->f64, d.stream);
        ctx->launches++;
        if (e == cudaSuccess && pos3) e = cudaMemcpyAsync(pos3 + 3 * d.shard_start, dp, cnt * 24, cudaMemcpyDeviceToHost, d.stream);
        if (e == cudaSuccess && vel3) e = cudaMemcpyAsync(vel3 + 3 * d.shard_start, dv, cnt * 24, cudaMemcpyDeviceToHost, d.stream);
        if (e == cudaSuccess && acc3) e = cudaMemcpyAsync(acc3 + 3 * d.shard_start, da, cnt * 24, cudaMemcpyDeviceToHost, d.stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(d.stream);
        cudaFree(scratch);
        CU(e);
    }
    return NBODY_OK;
}

int nbody_gpu_energy(nbody_ctx *ctx, double *K, double *W, double P[3])
{
    if (!ctx) return NBODY_EINVAL;
    int rc = sync_all(ctx);
    if (rc != NBODY_OK) return rc;
    const double eps2 = (double)ctx->p.eps * (double)ctx->p.eps;
    for (Dev &d : ctx->devs) {
        CU(cudaSetDevice(d.device));
        CU(cudaMemsetAsync(d.energy5, 0, 5 * sizeof(double), d.stream));
        CU(launch_energy(d.posm[d.cur], d.vel, ctx->n_padded, d.shard_start, d.shard_count, eps2, ctx->f64, d.energy5, d.stream));
        ctx->launches++;
    }
    if (ctx->p.world > 1) { // one process per GPU: sum over ranks on the device
        Dev &d = ctx->devs[0];
        CU(cudaSetDevice(d.device));
        NC(nccl().AllReduce(d.energy5, d.energy5, 5, NCCL_FLOAT64, NCCL_SUM, d.comm, d.stream));
    }
    double tot[5] = {0, 0, 0, 0, 0};
    for (Dev &d : ctx->devs) {
        double h[5];
        CU(cudaSetDevice(d.device));
        CU(cudaMemcpyAsync(h, d.energy5, sizeof h, cudaMemcpyDeviceToHost, d.stream));
        CU(cudaStreamSynchronize(d.stream));
        for (int k = 0; k < 5; ++k) tot[k] += h[k];
    }
    const double G = (double)ctx->p.G;
    if (K) *K = tot[0];
    if (W) *W = 0.5 * G * tot[1];
    if (P) { P[0] = tot[2]; P[1] = tot[3]; P[2] = tot[4]; }
    return NBODY_OK;
}

int nbody_gpu_profile_next_step(nbody_ctx *ctx, int enable)
{
    if (!ctx) return NBODY_EINVAL;
    ctx->profile_next = enable ? 1 : 0;
    return NBODY_OK;
}

int nbody_gpu_get_info(nbody_ctx *ctx, nbody_info *info)
{
    if (!ctx || !info) return NBODY_EINVAL;
    memset(info, 0, sizeof *info);
    const Dev &d0 = ctx->devs.front(), &dl = ctx->devs.back();
    info->n = ctx->n;
    info->n_padded = ctx->n_padded;
    info->shard_start = d0.shard_start;
    info->shard_count = dl.shard_start + dl.shard_count - d0.shard_start;
    info->world = ctx->world;
    info->rank = d0.rank;
    info->ngpus_local = (int)ctx->devs.size();
    info->p2p_exchange = ctx->p2p ? 1 : (ctx->ipc ? 2 : 0);
    info->sm_count = ctx->sm_count;
    info->sm_clock_khz = ctx->sm_clock_khz;
    info->j_splits = d0.plan.empty() ? 0 : d0.plan[0].splits;
    info->streamk_ctas = d0.plan.empty() ? 0u : (uint32_t)d0.plan[0].sk_ctas;
    info->force_ctas = d0.force_ctas;
    info->ctas_per_sm = ctx->ctas_per_sm;
    info->fused = d0.fused ? 1 : 0;
    info->uniform_mass = ctx->uniform ? 1 : 0;
    info->graph = (d0.graph != nullptr || d0.graph1[0] != nullptr || d0.graph1[1] != nullptr) ? 1 : 0;
    info->kernel_launches = ctx->launches;
    info->interactions = ctx->interactions;
    info->last_force_ms = ctx->last_force_ms;
    info->last_integ_ms = ctx->last_integ_ms;
    info->last_bh_build_ms = ctx->last_build_ms;
    info->last_collide_ms = ctx->last_collide_ms;
    info->last_bh_visits = ctx->last_visits;
    info->last_bh_visits_max = ctx->last_visits_max;
    if (ctx->bh) {
        unsigned m = 0;
        cudaSetDevice(ctx->devs[0].device);
        const cudaError_t ne = ctx->devs[0].bh.node_count(ctx->n, ctx->devs[0].stream, &m);
        info->bh_nodes = m;
        if (ne == cudaErrorMemoryAllocation) {
            set_err(ctx, "Barnes-Hut tree needs %u cells, %u reserved (raise NBODY_BH_NODE_FACTOR)", m, ctx->devs[0].bh.node_cap);
            return NBODY_ENOMEM;
        }
        if (ne != cudaSuccess) { set_err(ctx, "node_count: %s", cudaGetErrorString(ne)); return NBODY_ECUDA; }
    }
    return NBODY_OK;
}

int nbody_gpu_collide(nbody_ctx *ctx)
{
    if (!ctx) return NBODY_EINVAL;
    if (!ctx->p.collide) return NBODY_ESTATE;
    Dev &d = ctx->devs[0];
    CU(cudaSetDevice(d.device));
    int nl = 0;
    CU(d.col.run((float *)d.posm[d.cur], (float *)d.vel, ctx->n, d.stream, &nl));
    ctx->launches += (unsigned long long)nl;
    return NBODY_OK;
}

int nbody_gpu_collide_stats(nbody_ctx *ctx, uint32_t *candidate_pairs, uint32_t *resolved_pairs)
{
    if (!ctx) return NBODY_EINVAL;
    if (!ctx->p.collide) return NBODY_ESTATE;
    Dev &d = ctx->devs[0];
    CU(cudaSetDevice(d.device));
    unsigned c[4] = {0, 0, 0, 0};
    CU(d.col.stats(d.stream, c));
    if (c[2]) { set_err(ctx, "collision pass: grid entry / pair buffer overflow (radii far larger than the 600-unit cells?)"); return NBODY_ESTATE; }
    if (candidate_pairs) *candidate_pairs = c[1];
    if (resolved_pairs) *resolved_pairs = c[3];
    return NBODY_OK;
}

int nbody_gpu_bh_nodes(nbody_ctx *ctx, float *f8, uint32_t *u2, size_t cap, size_t *count)
{
    if (!ctx || !count) return NBODY_EINVAL;
    if (!ctx->bh) return NBODY_ESTATE;
    Dev &d = ctx->devs[0];
    CU(cudaSetDevice(d.device));
    unsigned m = 0;
    CU(d.bh.node_count(ctx->n, d.stream, &m));
    *count = m;
    if (cap == 0 || !f8 || !u2) return NBODY_OK;
    CU(d.bh.download_nodes(f8, u2, cap, d.stream));
    return NBODY_OK;
}

int nbody_gpu_streamk_owner(long long unit, long long units, int ctas)
{
    if (units <= 0 || ctas <= 0 || unit < 0 || unit >= units || (long long)ctas > units) return NBODY_EINVAL;
    return force_f32_streamk_owner(unit, units, ctas);
}

int nbody_gpu_streamk_slots(int tiles, int stages, int ctas)
{
    if (tiles <= 0 || stages <= 0 || ctas <= 0 || (long long)ctas > (long long)tiles * stages) return NBODY_EINVAL;
    return force_f32_streamk_slots(tiles, stages, ctas);
}

int nbody_gpu_nccl_unique_id(uint8_t id[NBODY_NCCL_ID_BYTES])
{
    if (!id) return NBODY_EINVAL;
    if (!nccl().load()) { set_err(nullptr, "libnccl.so.2 could not be loaded"); return NBODY_ENCCL; }
    ncclUniqueId u;
    if (nccl().GetUniqueId(&u) != 0) return NBODY_ENCCL;
    memcpy(id, &u, NBODY_NCCL_ID_BYTES);
    return NBODY_OK;
}

void nbody_gpu_shutdown(nbody_ctx *ctx)
{
    if (!ctx) return;
    if (ctx->ipc) {   // peers may still be storing into this rank's buffers: wait for them, then leave together
        if (sync_all(ctx) == NBODY_OK) ipc_barrier(ctx);
    }
    for (Dev &d : ctx->devs) {
        cudaSetDevice(d.device);
        if (d.comm_stream) cudaStreamSynchronize(d.comm_stream);
        if (d.stream) cudaStreamSynchronize(d.stream);
    }
    free_all(ctx);
}

const char *nbody_gpu_strerror(int code)
{
    switch (code) {
    case NBODY_OK: return "ok";
    case NBODY_EINVAL: return "invalid argument";
    case NBODY_ECUDA: return "CUDA error";
    case NBODY_ENOMEM: return "out of memory";
    case NBODY_ENCCL: return "NCCL error";
    case NBODY_ENODEV: return "no usable sm_100 device";
    case NBODY_ESTATE: return "invalid state";
    default: return "unknown error";
    }
}

const char *nbody_gpu_last_error(const nbody_ctx *ctx) { return ctx ? ctx->err : g_init_err; }

const char *nbody_gpu_version(void) { return "nbody_gpu 0.4 (sm_100a; all-pairs f32 fast[plain|uniform-mass]/refcompat, f64; Barnes-Hut quadtree/octree; peer exchange)"; }

} // extern "C"
