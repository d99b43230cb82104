/* nbody_gpu.h -- C ABI of the B200-native hot path of nbodysim.
 *
 * The reference (7IBBE77S/nbodysim) has no plugin / FFI interface: its hot path is the body of
 * `Simulation::step()` (Nbodysim/headers/Simulation.hpp:67-75), i.e.
 *     iterate(dt)  = attract()  [force accumulation, Simulation.hpp:176-214 -> Quadtree::acc,
 *                                Quadtree.hpp:113-155, per-pair kernel :136-143 / :119-127]
 *                  + kick / clamp / soft boundary / drift          [Simulation.hpp:129-163]
 * operating in place on the public `std::vector<Body> bodies` (Simulation.hpp:54).  This header
 * is the seam a maintainer binds instead: hand the `Body` array over once (`nbody_gpu_init`),
 * replace `iterate(current_dt)` by `nbody_gpu_step(ctx, current_dt, 1)`, and refresh the host
 * copy where the reference copies `simulation->bodies` for the renderer (main.cpp:623-627) with
 * `nbody_gpu_download`.  INTEGRATION.md shows the exact patch.
 *
 * Environment variables read at nbody_gpu_init (tuning / safety, never required):
 *   NBODY_PEER_TIMEOUT_S   seconds a rank waits for a peer's positions in the cross-process exchange before the
 *                          context fails with NBODY_ESTATE instead of spinning for ever (default 120)
 *   NBODY_BH_WALK_WINDOW   lane window of the warp-cooperative Barnes-Hut walk (default 256; 1 = lock-step lanes)
 *   NBODY_BH_NODE_FACTOR   Barnes-Hut cells reserved per body (default 4, at most 33)
 * A/B switches of the execution (read once per process; every setting gives the same bits, tests/test_gpu_bh.py):
 *   NBODY_SORT_LAZY=0      radix sort over all eight digits instead of the leading ones + run repair
 *   NBODY_SORT_COOP=0      one launch per digit pass instead of the all-passes cooperative kernel
 *   NBODY_BH_LOCAL=0       round-1 emit + climb-from-the-leaves instead of the window-local build
 *   NBODY_PDL=0            plain kernel launches instead of programmatic dependent launches along the Barnes-Hut step
 *   NBODY_BH_CTA_CLIMB=0   atomic climb over the top of the tree at every size (default: one CTA's shared memory up to 32,768 bodies)
 *   NBODY_BH_FUSE_INSERT=0 collision grid filled by its own kernel instead of the Barnes-Hut walk
 *   NBODY_COL_STRIP=W      width of the collision grid's x strips (default 37.5; 0 = whole cells)
 *   NBODY_BH_TRACE=1       per-phase clocks of the local build kernels on stderr (tuning; synchronises)
 *
 * Plain C: pointers and sizes only, no C++/torch types.  All functions return 0 on success or a
 * negative NBODY_E* code; they never throw, never call exit().  A context is used by one host
 * thread at a time (the reference drives step() from a single simulation thread, main.cpp:612-635).
 */
#ifndef NBODY_GPU_H
#define NBODY_GPU_H
#include "nbody_body.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct nbody_ctx nbody_ctx;

/* error codes */
#define NBODY_OK          0
#define NBODY_EINVAL     (-1)   /* bad argument */
#define NBODY_ECUDA      (-2)   /* CUDA runtime / driver error (see nbody_gpu_last_error) */
#define NBODY_ENOMEM     (-3)   /* device or host allocation failed */
#define NBODY_ENCCL      (-4)   /* NCCL missing or failed */
#define NBODY_ENODEV     (-5)   /* no usable sm_100 device */
#define NBODY_ESTATE     (-6)   /* call not valid in this state */

/* nbody_params.precision */
#define NBODY_PRECISION_F32 0
#define NBODY_PRECISION_F64 1   /* tolerance-check mode: state and arithmetic in double */
/* nbody_params.rsqrt_mode (fp32 only) */
#define NBODY_RSQRT_FAST      0 /* MUFU.RSQ (<= 2 ulp); packed FFMA2 pipeline; headline mode */
#define NBODY_RSQRT_REFCOMPAT 1 /* bit-faithful Quadtree::fast_inv_sqrt (Quadtree.hpp:106-111),
                                   unfused mul/add, source-order accumulation: reproduces the
                                   reference's direct sum bit for bit */
/* nbody_params.force_algo */
#define NBODY_FORCE_ALLPAIRS   0 /* direct sum == Quadtree::acc leaf loop, Quadtree.hpp:133-144 */
#define NBODY_FORCE_BARNES_HUT 1 /* tree walk == Quadtree::build + acc: the reference's quadtree for dims = 2,
                                    the same construction as an octree for dims = 3 */
/* nbody_params.integ_flags: extras of Simulation::iterate beyond Body::update */
#define NBODY_INTEG_CLAMP    1u /* |v| <= max_velocity            Simulation.hpp:133-137 */
#define NBODY_INTEG_BOUNDARY 2u /* exponential soft boundary + damping, Simulation.hpp:140-155 */
/* field masks for nbody_gpu_download */
#define NBODY_FIELD_POS 1u
#define NBODY_FIELD_VEL 2u
#define NBODY_FIELD_ACC 4u
#define NBODY_FIELD_ALL 7u

#define NBODY_MAX_GPUS 16
#define NBODY_NCCL_ID_BYTES 128

typedef struct nbody_params {
    uint32_t struct_size;     /* = sizeof(nbody_params); set by nbody_params_default */
    int32_t  dims;            /* 2 (reference) or 3 (z carried in the Vec2 padding) */
    float    eps;             /* Plummer softening length; reference ships 1.0 (Simulation.hpp:59) */
    float    G;               /* gravitational constant; reference is implicitly 1 */
    int32_t  precision;       /* NBODY_PRECISION_* */
    int32_t  rsqrt_mode;      /* NBODY_RSQRT_* */
    int32_t  force_algo;      /* NBODY_FORCE_* */
    float    theta;           /* Barnes-Hut opening angle; reference ships 1.0 */
    uint32_t integ_flags;     /* NBODY_INTEG_*; 0 == the clean kick-drift of Body::update */
    float    max_velocity;    /* 1000     Simulation.hpp:124 */
    float    boundary_radius; /* 100000   Simulation.hpp:120 */
    float    soft_boundary;   /* 0.8      Simulation.hpp:121 (fraction of boundary_radius) */
    float    boundary_force;  /* 0.9      Simulation.hpp:122 */
    float    damping;         /* 0.9995   Simulation.hpp:123 */
    int32_t  j_splits;        /* 0 = auto: the fast fp32 kernel runs in stream-K form (one persistent CTA per SM slot, equal runs of
                                 (target tile, source stage) units; at most a few partial slots per target); > 0 forces the split
                                 form with that many source-range splits (one partial slot per split) */
    int32_t  fuse_integrator; /* -1 = auto; 0 = separate integrator kernel; 1 (with j_splits = 1) = kick-drift fused
                                 into the force kernel's epilogue (2048-target tiles) */
    int32_t  use_graph;       /* -1 = auto (launch-bound sizes: n <= 32768, Barnes-Hut n <= 262144), 0 = never, 1 = always:
                                 replay pairs of steps from a CUDA graph in nbody_gpu_step calls of >= 8 steps (one GPU) */
    int32_t  force_variant;   /* fast fp32 kernel: -1 = auto, 0 = always the general-mass form (12 fp32
                                 lane-ops per interaction).  auto uses the uniform-mass form (11 lane-ops;
                                 m factored out of the sum, same per-pair arithmetic otherwise) when
                                 every body has the same mass. */
    int32_t  collide;         /* 1 = run the reference's collision pass after every step's integration
                                 (Simulation::collide + resolve, Simulation.hpp:216-346): Simulation::step()
                                 becomes nbody_gpu_step.  2-D fp32, one GPU.  Pairs are resolved in the
                                 canonical order "sorted by (first, second)"; the reference's order is the
                                 unspecified iteration order of an unordered_map, so results are bit-identical
                                 whenever the colliding pairs of a step are disjoint. */
    int32_t  bh_fix_near_leaves; /* Barnes-Hut only.  0 = the reference's behaviour: a NEAR leaf contributes
                                 nothing (insert() leaves every body Range empty, Quadtree.hpp:133-147);
                                 1 = add the leaf's body for near leaves (self excluded) */
    int32_t  sort_impl;       /* execution of the Barnes-Hut build and of the (rare) full collision pass.  0 = auto: the build runs
                                 one kernel per phase with the single-pass ("onesweep") radix sort and chained scan of
                                 csrc/radix_sort.cuh; the full collision pass of scenes of up to 65,536 bodies
                                 (NBODY_CLUSTER_MAX_N) runs as ONE CTA whose phases are separated by barriers, launched after the
                                 screening of every pass and returning at once when no two bodies sharing a grid cell overlap.
                                 1 = always one kernel per phase.  2 = also run the BUILD as one thread-block-cluster kernel
                                 (measured slower than 0 on B200, kept selectable; init fails if the device cannot host the
                                 cluster or n is too large).  Identical results. */
    int32_t  bh_walk;         /* Barnes-Hut only.  0 = auto, 1 = one independent walk per thread (targets in Z-order),
                                 2 = warp-cooperative walk (the warp walks the union of its 32 targets' traversals, every
                                 node record loaded once per warp).  Identical results bit for bit.  Measured: short
                                 walks (the reference's 2-D theta = 1, ~125 nodes) favour 1; long walks (3-D, or
                                 theta < 0.7) favour 2 by up to 1.8x -- auto picks accordingly. */
    int32_t  exchange;        /* how the new positions reach the other GPUs each step.  0 = auto: when every GPU can map
                                 every other one (ngpus > 1 in one process with peer access; or world > 1, one process
                                 per GPU on one node, peers mapped through CUDA IPC), the integrator kernel stores
                                 every new position directly into all peers' buffers over NVLink and publishes a
                                 completion counter there (integrate + allgather + signal in ONE kernel, no
                                 collective call); otherwise ncclAllGather on a communication stream.  1 = always
                                 NCCL.  2 = require the peer path (init fails instead of falling back). */
    /* --- single-process multi-GPU (C driver): ngpus devices, NCCL comms created internally --- */
    int32_t  ngpus;           /* 0 or 1 = single GPU */
    int32_t  device_ids[NBODY_MAX_GPUS]; /* CUDA ordinals; device_ids[0] is used when ngpus<=1 */
    /* --- one-process-per-GPU (torchrun): this process is `rank` of `world` --- */
    int32_t  world;           /* 0 or 1 = not distributed */
    int32_t  rank;
    uint8_t  nccl_id[NBODY_NCCL_ID_BYTES]; /* ncclUniqueId from nbody_gpu_nccl_unique_id on rank 0 */
    void    *stream;          /* optional cudaStream_t to launch on (single-GPU contexts); NULL = own */
} nbody_params;

typedef struct nbody_info {
    uint64_t n;               /* bodies given to init */
    uint64_t n_padded;        /* rounded up to whole target tiles per rank (zero-mass padding) */
    uint64_t shard_start;     /* first target owned by this process (all local GPUs) */
    uint64_t shard_count;
    int32_t  world, rank, ngpus_local;
    int32_t  p2p_exchange;    /* positions exchanged by peer stores from the integrator kernel: 1 = GPUs of one process,
                                 2 = across processes (CUDA IPC mappings + completion flags in peer memory); 0 = NCCL */
    int32_t  sm_count;        /* of the first local device */
    int32_t  sm_clock_khz;    /* cudaDevAttrClockRate */
    int32_t  j_splits;        /* partial-sum slots per target of the (first) force launch: the source-range splits of the split form,
                                 or the most CTAs that share a target tile in the stream-K form (see streamk_ctas) */
    int32_t  force_ctas;      /* CTAs per force launch (per GPU) */
    int32_t  ctas_per_sm;     /* resident force CTAs per SM (occupancy query) */
    int32_t  fused;           /* 1 if the kick-drift runs in the force kernel's epilogue */
    int32_t  uniform_mass;    /* 1 if the fast kernel runs the uniform-mass (11-op) form */
    int32_t  graph;           /* 1 if multi-step calls replay a CUDA graph */
    uint32_t bh_nodes;        /* Barnes-Hut: non-empty cells of the last tree built */
    uint32_t streamk_ctas;    /* > 0: the fast fp32 kernel runs its stream-K form on that many persistent CTAs */
    uint64_t kernel_launches; /* kernels of this library launched so far (all local GPUs) */
    uint64_t interactions;    /* pair interactions evaluated so far by this process */
    float    last_force_ms;   /* device time of the force kernel(s) of the last profiled step (Barnes-Hut: tree build + walk) */
    float    last_integ_ms;   /* device time of the integrator kernel of the last profiled step */
    float    last_bh_build_ms;/* Barnes-Hut: the tree-build part of last_force_ms (keys, sort, cells, centres of mass) */
    float    last_collide_ms; /* device time of the collision pass of the last profiled step (collide = 1) */
    uint64_t last_bh_visits;  /* Barnes-Hut: node records visited by the walk of the last profiled step, all targets */
    uint64_t last_bh_visits_max; /* ... and by the longest single walk (the per-thread walk kernel's critical path) */
} nbody_info;

/* Fill *p with the reference's shipped parameters (Simulation.hpp:59,120-124; G=1, dims=2,
 * all-pairs, fp32, fast rsqrt, clean integrator, one GPU = device 0). */
void nbody_params_default(nbody_params *p);

/* Create a context and upload `n` bodies (copied; caller keeps ownership of `bodies`).
 * Replaces: Simulation::Simulation() taking ownership of `bodies` (Simulation.hpp:58-65).
 * In distributed mode every rank passes the same full array. */
int nbody_gpu_init(nbody_ctx **out, const nbody_params *p, const nbody_body_t *bodies, size_t n);

/* Advance `nsteps` steps of size dt: per step, force accumulation on the current positions then
 * kick-drift (+ optional clamp/boundary).  Asynchronous: returns after enqueueing; the next
 * download/sync/energy call synchronises.  dt is per call because the reference re-reads
 * SIMULATION_DT every step (Simulation.hpp:69).
 * Replaces: Simulation::iterate(dt) called from Simulation::step() (Simulation.hpp:67-75). */
int nbody_gpu_step(nbody_ctx *ctx, float dt, int nsteps);

/* Evaluate accelerations at the current positions without integrating (Body::acc := a).
 * Replaces: Simulation::attract() (Simulation.hpp:176-214). */
int nbody_gpu_accel_only(nbody_ctx *ctx);

/* Block until all enqueued work of this context has finished. */
int nbody_gpu_sync(nbody_ctx *ctx);

/* Copy state back into the caller's `Body` array (first n records; only `fields` are written,
 * mass/radius are never modified).  Synchronises.  In distributed mode only this process's
 * shard [shard_start, shard_start+shard_count) is written.
 * Replaces: `SHARED_BODIES = simulation->bodies` (main.cpp:625). */
int nbody_gpu_download(nbody_ctx *ctx, nbody_body_t *bodies, size_t n, unsigned fields);

/* Replace the whole body state from host memory (positions, velocities, masses), e.g. after the
 * host ran Simulation::collide() on the downloaded bodies. n must equal the n given to init.
 * One process per GPU (world > 1): collective -- every rank calls it, and only this rank's shard
 * [shard_start, shard_start + shard_count) of `bodies` is read and copied to the device (the mirror image of
 * nbody_gpu_download); the packed shard reaches the other ranks over NVLink through the exchange. */
int nbody_gpu_upload(nbody_ctx *ctx, const nbody_body_t *bodies, size_t n);

/* Double-precision read-back of the shard (3 doubles per body each; any pointer may be NULL):
 * exact in NBODY_PRECISION_F64, widened fp32 otherwise.  For the tolerance checks. */
int nbody_gpu_download_f64(nbody_ctx *ctx, double *pos3, double *vel3, double *acc3, size_t n);

/* Diagnostics in fp64: kinetic energy K, potential W = -G sum_{i<j} m_i m_j /sqrt(r^2+eps^2)
 * (the potential consistent with the softened force), momentum P[3].  Synchronises.  New surface
 * (the reference has only the never-called Body::kinetic_energy/momentum, Body.hpp:98-106).
 * In distributed mode the values are already summed over all ranks. */
int nbody_gpu_energy(nbody_ctx *ctx, double *K, double *W, double P[3]);

/* Time the next nbody_gpu_step call's kernels with CUDA events (reported in nbody_info). */
int nbody_gpu_profile_next_step(nbody_ctx *ctx, int enable);

int nbody_gpu_get_info(nbody_ctx *ctx, nbody_info *info);

/* Run one collision pass on the current state (the reference's collide(), outside a step). */
int nbody_gpu_collide(nbody_ctx *ctx);
/* Counters of the last collision pass: candidate pairs kept after the broad phase and the pairs that
 * passed resolve()'s overlap test.  Synchronises. */
int nbody_gpu_collide_stats(nbody_ctx *ctx, uint32_t *candidate_pairs, uint32_t *resolved_pairs);

/* Barnes-Hut diagnostics / parity: the node array of the last tree built, in walk (depth-first,
 * quadrant) order, WITHOUT the reference's empty leaves.  f8 = 8 floats per node: position (body or
 * centre of mass) x,y,z ; mass ; cell centre x,y,z ; cell size  (Node::data, Node.hpp:35-40; z = 0 in 2-D).
 * u2 = 2 words per node: next (skip pointer, 0 = end of walk; Node::next) ; depth | is_leaf << 8.
 * Writes min(cap, count) nodes; *count receives the total. */
int nbody_gpu_bh_nodes(nbody_ctx *ctx, float *f8, uint32_t *u2, size_t cap, size_t *count);

/* Planning arithmetic of the stream-K force kernel, exposed for tests and capacity planning (pure host functions, no
 * device needed).  A launch is `tiles` target tiles x `stages` source stages = U units dealt to `ctas` persistent CTAs in
 * equal contiguous runs: unit u belongs to CTA nbody_gpu_streamk_owner(u, U, ctas); a tile's partial sums land in
 * owner(last unit of the tile) - owner(first unit) + 1 consecutive slots, and nbody_gpu_streamk_slots is the most any tile
 * needs (what the library reserves).  Requires 1 <= ctas <= tiles * stages. */
int nbody_gpu_streamk_owner(long long unit, long long units, int ctas);
int nbody_gpu_streamk_slots(int tiles, int stages, int ctas);

/* rank 0 creates the NCCL id that every rank passes in nbody_params.nccl_id. */
int nbody_gpu_nccl_unique_id(uint8_t id[NBODY_NCCL_ID_BYTES]);

void nbody_gpu_shutdown(nbody_ctx *ctx);

const char *nbody_gpu_strerror(int code);
/* Text of the last CUDA/NCCL failure seen by this context (or by init when ctx == NULL). */
const char *nbody_gpu_last_error(const nbody_ctx *ctx);
/* Library / build identification, e.g. "nbody_gpu 0.1 sm_100a". */
const char *nbody_gpu_version(void);

#ifdef __cplusplus
}
#endif
#endif /* NBODY_GPU_H */
