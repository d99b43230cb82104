// kernels.h -- launch interface between the C-ABI layer (nbody_gpu.cu) and the kernel files.
#pragma once
#include "common.cuh"
#include <algorithm>
#include <vector>

namespace nb {

// Integrator constants (defaults == Simulation.hpp:120-124), shared by the fused epilogue of the
// force kernel and the stand-alone integrator so both perform identical arithmetic.
struct IntegParams {
    float dt;
    float G;
    unsigned flags;          // NBODY_INTEG_*
    float max_velocity;      // MAX_VELOCITY
    float max_velocity_sq;   // MAX_VELOCITY * MAX_VELOCITY (fp32 product, as the reference forms it)
    float soft_boundary;     // BOUNDARY_RADIUS * 0.8f
    float soft_boundary_sq;  // SOFT_BOUNDARY * SOFT_BOUNDARY
    float boundary_force;    // BOUNDARY_FORCE
    float damping;           // DAMPING
};

// One force launch: targets = n_itiles tiles starting at block i_blk0; sources = blocks
// [j_blk0, j_blk0 + j_nblk) cut into `splits` chunks; CTA (tile, s) writes its partial sum into
// partial slot slot0 + s.  Slots are summed in index order by the integrator (deterministic).
struct ForceLaunch {
    int small_tile;          // fast kernel: 1 = 512-target tiles (small shards), 0 = 2048-target tiles
    int dims;                // fast kernel: 2 = planar data (z == 0 everywhere): z operations dropped
    int uniform_mass;        // fast kernel: 1 = every massive source has the same mass (11-op form)
    float acc_scale;         // fused epilogue: G (plain) or G*m (uniform)
    const void *posm;        // blocked (x,y,z,m), float or double
    void *accp;              // partial accelerations, blocked, [slot][local block]
    int i_blk0;              // first target block (global block index into posm)
    int i_blk_local0;        // same block's index inside this GPU's shard (for accp / vel / acc)
    int n_iblk;              // target blocks in this launch
    int n_iblk_shard;        // blocks in the whole shard (slot stride of accp)
    int j_blk0, j_nblk;      // source block range
    long long j_body_limit;  // sources >= this global body index are padding (refcompat skips them)
    int splits, slot0;
    int streamk_ctas;        // fast kernel: > 0 = stream-K form with that many persistent CTAs (splits ignored); 0 = split form
    float eps2;
    double eps2_f64;
    // fused kick-drift epilogue (fast fp32 kernel, splits == 1, single launch per step)
    int fuse;
    void *posm_next;         // blocked, full array (only the target blocks are written)
    void *vel;               // blocked, shard-local
    void *acc;               // blocked, shard-local
    IntegParams ip;
};

// Destinations of the integrator's new positions.  One entry (this GPU's next buffer) normally; with the
// peer-to-peer exchange, the next buffer of EVERY GPU of the process: the kick-drift kernel then stores each
// new position straight into all peers' memory over NVLink -- integrate and allgather in one kernel.
struct PeerDests {
    void *p[16];
    int n;
};

// Which partial slots hold the sum of a target block.  A step's force launches each own a run of slots:
//   split launch      `nslots` slots, all filled for every target
//   stream-K launch   the slots of target tile T are slot0 .. slot0 + owner(T S + S - 1) - owner(T S), with
//                     owner(u) = ((u + 1) G - 1) / U  (force_f32_fast.cuh)
struct SlotRange {
    int slot0, nslots;
    int sk_S, sk_G, tile_blks;   // stream-K: source stages per tile (0 = split launch), CTAs, target blocks per tile
    long long sk_U;              // stream-K: units of the launch (tiles x stages)
};
struct SlotPlan {
    int n;
    SlotRange r[4];
};

struct IntegLaunch {
    const void *posm_cur;    // blocked, full array
    void *posm_next;         // blocked, full array
    PeerDests dests;         // where the new positions are stored (n >= 1; dests.p[0] == posm_next unless peer exchange)
    PeerSignal signal;       // cross-process exchange: completion counters to publish (n == 0: none)
    void *vel, *acc;         // blocked, shard-local
    const void *accp;        // partial slots
    float acc_scale;         // G, or G*m when the uniform-mass force kernel summed unit masses
    SlotPlan slots;
    int i_blk0;              // first global block of the shard
    int n_iblk_shard;
    int acc_only;            // 1: acc := G * sum(partials), no kick-drift (nbody_gpu_accel_only)
    long long n_real;        // bodies >= n_real are zero-mass padding: never moved
    IntegParams ip;
};

// geometry of each force kernel variant (fast: chosen by the tools/kbench.cu sweeps on B200 --
// 8 targets per thread, 256 threads, one CTA of 8 warps per SM, ~228 registers per thread)
constexpr int FAST_THREADS = 256, FAST_I = 8, FAST_MINB = 1, FAST_UNROLL = 1, FAST_STAGE_BLKS = 2;
constexpr int FAST_TILE_BLKS = FAST_I / (BLK / FAST_THREADS);                // 8 blocks = 2048 targets / CTA
// small shards (the reference's own N = 25,000 included): 512-target tiles so that the grid still covers
// the 148 SMs; 4 targets per thread, 128 threads, 4 CTAs per SM, one source block per stage (kbench sweep)
constexpr int SMALL_THREADS = 128, SMALL_I = 4, SMALL_MINB = 4, SMALL_UNROLL = 2, SMALL_STAGE_BLKS = 1;
constexpr int SMALL_TILE_BLKS = SMALL_I / (BLK / SMALL_THREADS);             // 2 blocks = 512 targets / CTA
constexpr int REF_THREADS = 128, REF_TILE_BODIES = 128;                      // refcompat: 1 / thread
constexpr int F64_THREADS = 128, F64_I = 2, F64_TILE_BLKS = 1;               // 256 targets / CTA
constexpr int TARGET_GRANULE = FAST_TILE_BLKS * BLK;                         // shard granularity

cudaError_t launch_force_f32_fast(const ForceLaunch &L, bool guard_zero, cudaStream_t st);
cudaError_t launch_force_f32_refcompat(const ForceLaunch &L, cudaStream_t st);
cudaError_t launch_force_f64(const ForceLaunch &L, cudaStream_t st);
int force_f32_fast_ctas_per_sm(bool uniform_mass, bool small_tile = false);
int force_f32_streamk_ctas_per_sm(bool uniform_mass, bool small_tile);
// slots a stream-K launch of `tiles` target tiles x `stages` source stages on `ctas` CTAs needs (the most any tile uses)
int force_f32_streamk_slots(int tiles, int stages, int ctas);
int force_f32_streamk_owner(long long unit, long long units, int ctas);   // CTA that owns a unit (sk_owner)
int force_f32_fast_grid(const ForceLaunch &L);

cudaError_t launch_integrate_f32(const IntegLaunch &L, cudaStream_t st);
cudaError_t launch_integrate_f64(const IntegLaunch &L, cudaStream_t st);
// cross-process exchange: block the stream until every peer's counter in `flags[0..world)` (this rank's own
// flag array) has reached `need`; after `timeout_ns` without progress *status is set to 1 and the kernel ends
cudaError_t launch_wait_peer_flags(const unsigned long long *flags, int world, int self, unsigned long long need,
                                   unsigned long long timeout_ns, unsigned *status, cudaStream_t st);

// AoS (reference Body, 64 B) <-> blocked SoA
// packs records [first, first + count) (global body indices; `aos` starts at record aos_first of the caller's array)
cudaError_t launch_pack(const void *aos, size_t aos_first, size_t first, size_t count, size_t n, size_t shard_start,
                        size_t shard_count, void *posm, void *vel, void *acc, bool f64, int dims,
                        float check_mass, unsigned *status, cudaStream_t st);
cudaError_t launch_unpack(void *aos, size_t n, size_t shard_start, size_t shard_count,
                          const void *posm, const void *vel, const void *acc, bool f64,
                          cudaStream_t st);
cudaError_t launch_unpack_f64(double *pos3, double *vel3, double *acc3, size_t n,
                              size_t shard_start, size_t shard_count, const void *posm,
                              const void *vel, const void *acc, bool f64, cudaStream_t st);

// fp64 diagnostics: out[0]=K, out[1]=W (pairs counted twice, caller halves), out[2..4]=P
cudaError_t launch_energy(const void *posm, const void *vel, size_t n_padded, size_t shard_start,
                          size_t shard_count, double eps2, bool f64, double *out5, cudaStream_t st);

// ---- collision pass (collide.cu / collide.cuh): the arguments of a pass and its screening hash grid
// counters: [0] cell entries, [1] pairs kept (hot components), [2] overflow flag, [3] pairs resolved (narrow test
// passed), [4] sweep pairs that overlap now
struct ColArgs {
    float *posm, *vel;                     // blocked SoA; radius rides in vel's 4th component
    unsigned n;
    unsigned long long *keys_in;           // cell entries as produced (hash, body) ...
    unsigned *vals_in;
    const unsigned long long *keys;        // ... and sorted by hash
    const unsigned *vals;
    unsigned entry_cap;
    unsigned long long *pairs;             // pair keys as produced
    unsigned pair_cap;
    unsigned char *hot;                    // [0, n): body is in a pair that overlaps now ; [n, 2n): component label is hot
    unsigned *parent;                      // union-find forest over the bodies
    unsigned *counters;
    unsigned *status;                      // host-visible sticky flags ([1] = collision buffers overflowed), may be null
    const unsigned *gate;                  // screening result ([0] = pairs sharing a cell that overlap now); 0 there: nothing to do
    int idx_bits, rooted;
    float strip;                           // width of the x strips the screening grid splits a cell into (0: none)
};

// open-addressing table keyed by (cell, x strip), one linked list of bodies per key (collide.cuh)
struct ColGrid {
    unsigned long long *tkeys;             // [tmask + 1]  0 = empty, else 1 << 32 | cell hash
    unsigned *heads;                       // [tmask + 1]  entry index + 1 of the cell's list head, 0 = none
    unsigned *enext;                       // [ecap]       list links (entry index + 1)
    float4 *edata;                         // [ecap]       (x, y, radius, body index bits) of the entry's body: one load per list element
    unsigned *flags;                       // [0] overlapping pairs seen, [1] entries used       (zeroed with the table)
    unsigned tmask, ecap;
};

// fused kick-drift epilogue of the Barnes-Hut walk (small scenes on one GPU): where the walk threads find velocities
// and store the new state
struct BhFuseArgs {
    float *posm_next, *vel, *acc;
    float G;
    IntegParams ip;
    // collision pass follows: the walk threads also enter their (new) positions into its screening grid
    const ColArgs *col_args = nullptr;
    const ColGrid *col_grid = nullptr;
};

// Barnes-Hut path (barnes_hut.cu): per-GPU workspace holding sorted keys and the pre-order node array
struct BhWorkspace {
    size_t n_cap = 0;
    unsigned node_cap = 0, n_nodes = 0;
    void *root = nullptr, *box = nullptr, *keys_in = nullptr, *keys = nullptr, *idx_in = nullptr, *idx = nullptr;
    void *count = nullptr, *offs = nullptr, *first = nullptr, *leaf = nullptr;
    void *node_data = nullptr, *node_quad = nullptr, *node_arrive = nullptr, *node_slots = nullptr;   // node_data: one 32-byte record per node
    void *trace = nullptr;       // tuning aid (NBODY_CLUSTER_TRACE): per-phase clock stamps of the cluster build
    void *node_slot_cells = nullptr; // centre-of-mass pass: subtree sizes riding up with the centres of mass (-> skip pointers)
    void *shard_targets = nullptr; // several GPUs: this GPU's targets compacted out of the Z-order (+ their count)
    void *node_owner = nullptr;  // single-cluster build: the sorted body that owns each cell
    int top_cta_state = 0;       // one-CTA top-of-tree kernel: 0 = not set up on this device yet, 1 = usable, -1 = not
    void *climb_start = nullptr; // per sorted body: the cell where its chain leaves the CTA-local part of the centre-of-mass pass
    int cluster_ctas = 0;        // > 0: scenes of up to cluster_ctas x 49152 bodies are built by ONE cluster kernel of that many CTAs
    static int cluster_ctas_available(int dims);
    int dims = 2;                // 2 = the reference's quadtree, 3 = octree
    unsigned *status = nullptr;  // host-visible sticky flags ([0] = more cells than reserved), owned by the context
    // cleared by one memset at the start of every build: box | sort scratch | scan scratch | arrival counters
    void *zero_region = nullptr, *sort_temp = nullptr, *scan_temp = nullptr, *shard_scan = nullptr;
    size_t zero_bytes = 0;
    bool count_valid = false;
    bool warp_walk = false;      // warp-cooperative walk, or (default) one independent walk per thread
    unsigned walk_window = 256;  // warp-cooperative walk: how many records ahead of the slowest lane a lane may run
                                 // (sweep on B200, tools/bh_window_sweep.py: 256 is best or within 1 % of best everywhere)
    cudaError_t alloc(size_t n, int dims, double node_factor, int cluster_mode, size_t cluster_max_n);
    cudaError_t node_count(size_t n, cudaStream_t st, unsigned *out);
    void release();
    cudaError_t build(const float *posm, size_t n, cudaStream_t st, int *launches);
    cudaError_t walk(const float *posm, size_t n, float theta, float eps, bool refcompat, bool fix_near_leaves,
                     size_t shard_start, size_t shard_count, float *accp, unsigned long long *visits,
                     const BhFuseArgs *fuse, cudaStream_t st);
    cudaError_t download_nodes(float *f8, unsigned *u2, size_t cap, cudaStream_t st);
};

// collision pass (collide.cu)
struct CollideWorkspace {
    size_t n_cap = 0;
    unsigned entry_cap = 0, pair_cap = 0;
    void *keys_in = nullptr, *keys = nullptr, *vals_in = nullptr, *vals = nullptr;
    void *pairs_in = nullptr, *pairs = nullptr, *hot = nullptr, *parent = nullptr, *counters = nullptr, *temp = nullptr;
    size_t temp_bytes = 0;
    unsigned *status = nullptr;  // host-visible sticky flags ([1] = entry / pair buffers overflowed), owned by the context
    void *grid = nullptr, *grid_links = nullptr, *ranks = nullptr;   // screening hash grid (flags | keys | heads), its list links; sort ranks
    size_t grid_bytes = 0;
    unsigned table_slots = 0;
    bool single_cta = false;     // small scenes: the (rare) full pass runs as ONE CTA, phases separated by barriers
    cudaError_t alloc(size_t n, int single_cta_mode, size_t single_cta_max_n);
    void release();
    ColArgs args(float *posm, float *vel, size_t n) const;
    ColGrid grid_view() const;
    cudaError_t prepare(cudaStream_t st);    // clears the screening grid (before whoever fills it)
    // `grid_filled`: prepare() was called and the bodies are in the grid already (the fused Barnes-Hut walk entered them)
    cudaError_t run(float *posm, float *vel, size_t n, cudaStream_t st, int *launches, bool grid_filled = false);
    cudaError_t stats(cudaStream_t st, unsigned out[4]);
};

} // namespace nb
