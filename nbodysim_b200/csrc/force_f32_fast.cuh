// force_f32_fast.cuh -- the headline all-pairs kernel (template), shared by the product library
// (force_f32.cu) and the tuning harness (tools/kbench.cu).  See force_f32.cu for the design notes.
#pragma once
#include "kernels.h"

namespace nb {

constexpr int NSTAGE = 4;                           // ring depth of the TMA source pipeline

// ---- the one place the integrator arithmetic lives (device side, fp32) -----------------------
// Body::update (Body.hpp:34-38): vel += acc*dt ; pos += vel*dt, as unfused mul-then-add (two
// roundings each) exactly like the strict build of the reference; plus the optional extras of
// Simulation::iterate (Simulation.hpp:129-155).
__device__ __forceinline__ void integrate_body_f32(float &px, float &py, float &pz, float &vx,
                                                   float &vy, float &vz, float ax, float ay,
                                                   float az, const IntegParams &ip)
{
    const float dt = ip.dt;
    vx = __fadd_rn(vx, __fmul_rn(ax, dt));
    vy = __fadd_rn(vy, __fmul_rn(ay, dt));
    vz = __fadd_rn(vz, __fmul_rn(az, dt));
    if (ip.flags & 1u) { // Simulation.hpp:133-137
        float v2 = __fadd_rn(__fadd_rn(__fmul_rn(vx, vx), __fmul_rn(vy, vy)), __fmul_rn(vz, vz));
        if (v2 > ip.max_velocity_sq) {
            float scale = __fdiv_rn(ip.max_velocity, __fsqrt_rn(v2));
            vx = __fmul_rn(vx, scale);
            vy = __fmul_rn(vy, scale);
            vz = __fmul_rn(vz, scale);
        }
    }
    if (ip.flags & 2u) { // Simulation.hpp:142-155
        float d2 = __fadd_rn(__fadd_rn(__fmul_rn(px, px), __fmul_rn(py, py)), __fmul_rn(pz, pz));
        if (d2 > ip.soft_boundary_sq) {
            float dist = __fsqrt_rn(d2);
            float ratio = __fdiv_rn(dist, ip.soft_boundary);
            float force = __fmul_rn(ip.boundary_force, expf(__fsub_rn(ratio, 1.0f)));
            float k = __fdiv_rn(-1.0f, dist);
            float fd = __fmul_rn(force, dt);
            vx = __fadd_rn(vx, __fmul_rn(__fmul_rn(px, k), fd));
            vy = __fadd_rn(vy, __fmul_rn(__fmul_rn(py, k), fd));
            vz = __fadd_rn(vz, __fmul_rn(__fmul_rn(pz, k), fd));
            vx = __fmul_rn(vx, ip.damping);
            vy = __fmul_rn(vy, ip.damping);
            vz = __fmul_rn(vz, ip.damping);
        }
    }
    px = __fadd_rn(px, __fmul_rn(vx, dt));
    py = __fadd_rn(py, __fmul_rn(vy, dt));
    pz = __fadd_rn(pz, __fmul_rn(vz, dt));
}

// ---- TMA source-tile ring -----------------------------------------------------------------------
// SRC_BLK_ELEMS floats per source block (4 or 5 component arrays of 256), STAGE_BLKS blocks per stage.
template <int SRC_BLK_ELEMS, int STAGE_BLKS>
struct Ring {
    static constexpr int STAGE_FLOATS = STAGE_BLKS * SRC_BLK_ELEMS;
    static constexpr int STAGE_BYTES = STAGE_FLOATS * 4;
    static constexpr size_t SMEM = (size_t)NSTAGE * STAGE_BYTES + 2 * NSTAGE * sizeof(uint64_t);
    float *stage;        // NSTAGE * STAGE_FLOATS
    uint64_t *full;      // NSTAGE, tx-count barriers armed by the producer thread
    uint64_t *empty;     // NSTAGE, one arrival per consumer warp

    __device__ __forceinline__ void setup(unsigned char *smem_raw, int nwarps)
    {
        stage = reinterpret_cast<float *>(smem_raw);
        full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)NSTAGE * STAGE_BYTES);
        empty = full + NSTAGE;
        if (threadIdx.x == 0) {
            for (int s = 0; s < NSTAGE; ++s) {
                mbar_init(&full[s], 1);
                mbar_init(&empty[s], nwarps);
            }
            mbar_fence_init();
        }
        __syncthreads();
    }
    __device__ __forceinline__ void issue(const float *src_blocks, int t, int chunk_blks) const
    {
        const int s = t % NSTAGE;
        const int nb = min(STAGE_BLKS, chunk_blks - t * STAGE_BLKS);
        const uint32_t bytes = (uint32_t)nb * SRC_BLK_ELEMS * 4u;
        mbar_expect_tx(&full[s], bytes);
        tma_bulk_g2s(stage + (size_t)s * STAGE_FLOATS,
                     src_blocks + (size_t)t * STAGE_BLKS * SRC_BLK_ELEMS, bytes, &full[s]);
    }
    // stage buffer s <- `bytes` bytes at gsrc (whole source blocks)
    __device__ __forceinline__ void issue_raw(int s, const float *gsrc, uint32_t bytes) const
    {
        mbar_expect_tx(&full[s], bytes);
        tma_bulk_g2s(stage + (size_t)s * STAGE_FLOATS, gsrc, bytes, &full[s]);
    }
    // consumer side of stage t is done; thread 0 refills the buffer of the PREVIOUS stage
    __device__ __forceinline__ void release_and_refill(const float *src_blocks, int t, int nst,
                                                       int chunk_blks) const
    {
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[t % NSTAGE]);
        if (threadIdx.x == 0 && t >= 1 && (t - 1 + NSTAGE) < nst) {
            const int tp = t - 1;
            mbar_wait(&empty[tp % NSTAGE], (uint32_t)(tp / NSTAGE) & 1u);
            issue(src_blocks, tp + NSTAGE, chunk_blks);
        }
    }
};

__device__ __forceinline__ float2 lo2(const float4 v) { return make_float2(v.x, v.y); }
__device__ __forceinline__ float2 hi2(const float4 v) { return make_float2(v.z, v.w); }

// Plain formulation, 12 fp32-pipe lane-ops + 1 MUFU per interaction:
//   d = p_j - p_i (3 FADD) ; r2 = d.d + eps^2 (3 FFMA) ; ri = rsqrt(r2) ; s = m_j ri^3 (3 FMUL) ;
//   acc += d s (3 FFMA)
// DIMS = 2 drops every z operation (the reference itself is 2-D): 9 / 8 lane-ops instead of 12 / 11.
template <int I, bool GUARD, int DIMS>
__device__ __forceinline__ void interact_pair_plain(const float2 xj, const float2 yj,
                                                    const float2 zj, const float2 mj,
                                                    const float2 (&nxi)[I], const float2 (&nyi)[I],
                                                    const float2 (&nzi)[I], float2 (&ax)[I],
                                                    float2 (&ay)[I], float2 (&az)[I], const float2 e2)
{
#pragma unroll
    for (int k = 0; k < I; ++k) {
        const float2 dx = __fadd2_rn(xj, nxi[k]);
        const float2 dy = __fadd2_rn(yj, nyi[k]);
        const float2 dz = (DIMS == 3) ? __fadd2_rn(zj, nzi[k]) : make_float2(0.f, 0.f);
        float2 r2 = __ffma2_rn(dx, dx, e2);
        r2 = __ffma2_rn(dy, dy, r2);
        if (DIMS == 3) r2 = __ffma2_rn(dz, dz, r2);
        float2 ri = make_float2(rsqrt_approx(r2.x), rsqrt_approx(r2.y));
        if (GUARD) { // eps == 0: self / coincident pairs contribute nothing (Quadtree.hpp:139)
            ri.x = (r2.x > 0.0f) ? ri.x : 0.0f;
            ri.y = (r2.y > 0.0f) ? ri.y : 0.0f;
        }
        const float2 ri2 = __fmul2_rn(ri, ri);
        const float2 mr = __fmul2_rn(mj, ri);
        const float2 s = __fmul2_rn(mr, ri2);
        ax[k] = __ffma2_rn(dx, s, ax[k]);
        ay[k] = __ffma2_rn(dy, s, ay[k]);
        if (DIMS == 3) az[k] = __ffma2_rn(dz, s, az[k]);
    }
}

// Uniform-mass formulation, 11 lane-ops + 1 MUFU, exact displacements: when every massive source has
// the same mass the multiply by m_j factors out of the sum (applied once per target afterwards):
//   d = p_j - p_i (3 FADD) ; r2 = d.d + eps^2 (3 FFMA) ; ri = rsqrt(r2) ; s = ri^3 (2 FMUL) ; acc += d s
// Padding sources sit at 1e18 (pack kernel): ri^3 underflows to exactly 0 there.
template <int I, bool GUARD, int DIMS>
__device__ __forceinline__ void interact_pair_uniform(const float2 xj, const float2 yj,
                                                      const float2 zj, const float2 (&nxi)[I],
                                                      const float2 (&nyi)[I], const float2 (&nzi)[I],
                                                      float2 (&ax)[I], float2 (&ay)[I], float2 (&az)[I],
                                                      const float2 e2)
{
#pragma unroll
    for (int k = 0; k < I; ++k) {
        const float2 dx = __fadd2_rn(xj, nxi[k]);
        const float2 dy = __fadd2_rn(yj, nyi[k]);
        const float2 dz = (DIMS == 3) ? __fadd2_rn(zj, nzi[k]) : make_float2(0.f, 0.f);
        float2 r2 = __ffma2_rn(dx, dx, e2);
        r2 = __ffma2_rn(dy, dy, r2);
        if (DIMS == 3) r2 = __ffma2_rn(dz, dz, r2);
        float2 ri = make_float2(rsqrt_approx(r2.x), rsqrt_approx(r2.y));
        if (GUARD) {
            ri.x = (r2.x > 0.0f) ? ri.x : 0.0f;
            ri.y = (r2.y > 0.0f) ? ri.y : 0.0f;
        }
        const float2 ri2 = __fmul2_rn(ri, ri);
        const float2 s = __fmul2_rn(ri2, ri);
        ax[k] = __ffma2_rn(dx, s, ax[k]);
        ay[k] = __ffma2_rn(dy, s, ay[k]);
        if (DIMS == 3) az[k] = __ffma2_rn(dz, s, az[k]);
    }
}

enum { FORM_PLAIN = 0, FORM_UNIFORM = 1 };

struct FastArgs {
    const float *posm;       // blocked (x,y,z,m): sources and targets
    float *accp;
    int i_blk0, i_blk_local0, n_iblk_shard, j_blk0, j_nblk, splits, slot0;
    int n_tiles;             // stream-K kernel: target tiles of the launch
    float eps2;
    float acc_scale;         // fused epilogue factor: G (plain) or G*m (uniform)
    long long n_real;
    float *posm_next, *vel, *acc;   // fused epilogue only
    IntegParams ip;
};

template <int I, int THREADS, int MINB, int UNROLL, int STAGE_BLKS, int FORM, bool GUARD, bool FUSE, int DIMS = 3>
__global__ void __launch_bounds__(THREADS, MINB) force_f32_fast_kernel(const FastArgs a)
{
    constexpr int SRC_ELEMS = BLK_ELEMS;
    using RingT = Ring<SRC_ELEMS, STAGE_BLKS>;
    constexpr int LANES_PER_BLK = BLK / THREADS;          // 1 (256 threads) or 2 (128 threads)
    static_assert(BLK % THREADS == 0 && I % LANES_PER_BLK == 0, "tile shape");
    constexpr int TILE_BLKS = I / LANES_PER_BLK;           // target blocks per CTA
    extern __shared__ __align__(128) unsigned char smem_raw[];
    RingT ring;
    ring.setup(smem_raw, THREADS / 32);

    const int tile = blockIdx.x / a.splits;
    const int split = blockIdx.x - tile * a.splits;
    const int tid = threadIdx.x;

    // source chunk of this CTA: whole blocks, balanced to within one block
    const int jb0 = a.j_blk0 + (int)(((long long)a.j_nblk * split) / a.splits);
    const int jb1 = a.j_blk0 + (int)(((long long)a.j_nblk * (split + 1)) / a.splits);
    const int chunk_blks = jb1 - jb0;
    const int nst = (chunk_blks + STAGE_BLKS - 1) / STAGE_BLKS;
    const float *src = a.posm + (size_t)jb0 * SRC_ELEMS;

    if (tid == 0) {
        const int pre = min(NSTAGE, nst);
        for (int t = 0; t < pre; ++t) ring.issue(src, t, chunk_blks);
    }

    // targets: I bodies per thread, coalesced block reads.  Keep -p_i broadcast over both packed
    // lanes so that r = p_j + (-p_i) is one FADD2 (ptxas folds the pair into a .F32 operand).
    float2 nxi[I], nyi[I], nzi[I], ax[I], ay[I], az[I];
    size_t tgt_off[I];
#pragma unroll
    for (int k = 0; k < I; ++k) {
        const int blk = tile * TILE_BLKS + k / LANES_PER_BLK;
        const int lane = (k % LANES_PER_BLK) * THREADS + tid;
        tgt_off[k] = (size_t)blk * BLK_ELEMS + lane;
        const float *b = a.posm + (size_t)a.i_blk0 * BLK_ELEMS + tgt_off[k];
        const float x = b[0], y = b[BLK], z = b[2 * BLK];
        nxi[k] = make_float2(-x, -x);
        nyi[k] = make_float2(-y, -y);
        nzi[k] = make_float2(-z, -z);
        ax[k] = ay[k] = az[k] = make_float2(0.f, 0.f);
    }
    const float2 e2 = make_float2(a.eps2, a.eps2);

    for (int t = 0; t < nst; ++t) {
        const int s = t % NSTAGE;
        mbar_wait(&ring.full[s], (uint32_t)(t / NSTAGE) & 1u);
        const float *st = ring.stage + (size_t)s * RingT::STAGE_FLOATS;
        const int nb = min(STAGE_BLKS, chunk_blks - t * STAGE_BLKS);
        for (int b = 0; b < nb; ++b) {
            const float *sx = st + b * SRC_ELEMS;
#pragma unroll UNROLL
            for (int j = 0; j < BLK; j += 4) {
                const float4 X = *reinterpret_cast<const float4 *>(sx + j);
                const float4 Y = *reinterpret_cast<const float4 *>(sx + BLK + j);
                const float4 Z = (DIMS == 3) ? *reinterpret_cast<const float4 *>(sx + 2 * BLK + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                if (FORM == FORM_UNIFORM) {
                    interact_pair_uniform<I, GUARD, DIMS>(lo2(X), lo2(Y), lo2(Z), nxi, nyi, nzi, ax, ay, az, e2);
                    interact_pair_uniform<I, GUARD, DIMS>(hi2(X), hi2(Y), hi2(Z), nxi, nyi, nzi, ax, ay, az, e2);
                    continue;
                }
                const float4 M = *reinterpret_cast<const float4 *>(sx + 3 * BLK + j);
                interact_pair_plain<I, GUARD, DIMS>(lo2(X), lo2(Y), lo2(Z), lo2(M), nxi, nyi, nzi, ax, ay, az, e2);
                interact_pair_plain<I, GUARD, DIMS>(hi2(X), hi2(Y), hi2(Z), hi2(M), nxi, nyi, nzi, ax, ay, az, e2);
            }
        }
        ring.release_and_refill(src, t, nst, chunk_blks);
    }

    // epilogue: fold the two packed lanes (even/odd sources)
#pragma unroll
    for (int k = 0; k < I; ++k) {
        const float fx = ax[k].x + ax[k].y, fy = ay[k].x + ay[k].y, fz = az[k].x + az[k].y;
        const size_t loff = (size_t)a.i_blk_local0 * BLK_ELEMS + tgt_off[k];   // inside the shard
        if (FUSE) {
            const size_t goff = (size_t)a.i_blk0 * BLK_ELEMS + tgt_off[k];
            const long long body = (long long)(goff / BLK_ELEMS) * BLK + (long long)(goff % BLK);
            if (body >= a.n_real) continue;                      // padding stays put
            // fused kick-drift: the new positions go to the other posm buffer, so CTAs still
            // reading the current one are undisturbed (race-free by construction).
            const float g = a.acc_scale;
            const float gx = fx * g, gy = fy * g, gz = fz * g;
            const float *pb = a.posm + goff;
            float *vb = a.vel + loff, *ab = a.acc + loff, *nb_ = a.posm_next + goff;
            float px = pb[0], py = pb[BLK], pz = pb[2 * BLK];
            const float m = pb[3 * BLK];
            float vx = vb[0], vy = vb[BLK], vz = vb[2 * BLK];
            integrate_body_f32(px, py, pz, vx, vy, vz, gx, gy, gz, a.ip);
            nb_[0] = px; nb_[BLK] = py; nb_[2 * BLK] = pz; nb_[3 * BLK] = m;
            vb[0] = vx; vb[BLK] = vy; vb[2 * BLK] = vz;
            ab[0] = gx; ab[BLK] = gy; ab[2 * BLK] = gz;
        } else {
            float *o = a.accp + (size_t)(a.slot0 + split) * a.n_iblk_shard * BLK_ELEMS + loff;
            o[0] = fx; o[BLK] = fy; o[2 * BLK] = fz;
        }
    }
}

// ---- stream-K form of the same kernel ---------------------------------------------------------------------------------------
// The work of a launch is (target tiles) x (source stages).  The split form above cuts it into tiles x splits CTAs and
// every CTA writes one partial sum per target -- 13 partial slots at N = 1M, 116 MB written by the force kernel and
// read back by the integrator.  Here the launch is ONE persistent CTA per SM slot, and the tile-major sequence of
// (tile, stage) units is dealt out in equal contiguous runs: CTA k takes units [k U / G, (k+1) U / G).  Inside a run the
// TMA ring never drains -- the source stage of a unit does not depend on the tile, so crossing a tile boundary only
// means storing the finished partial sum and loading the next tile's targets while the ring keeps prefetching.  A tile
// is shared by at most (stages per tile) / (units per CTA) + 2 CTAs -- two at N = 1M -- each writing its own slot, in
// unit order; the integrator derives the number of slots of a tile from the same arithmetic and sums them in that
// order (deterministic).  Balance: every CTA gets the same number of units to within one stage (512 sources).
constexpr int SK_FLUSH = 8;                        // stages (4,096 sources with 2-block stages) between folds of the register sums
__host__ __device__ constexpr size_t SK_SACC_OFFSET(size_t ring_bytes) { return (ring_bytes + 15) & ~(size_t)15; }
__host__ __device__ constexpr size_t SK_SMEM(size_t ring_bytes, int targets_per_thread, int threads)
{
    return SK_SACC_OFFSET(ring_bytes) + (size_t)3 * targets_per_thread * threads * sizeof(float);
}
__host__ __device__ __forceinline__ long long sk_start(long long k, long long U, int G) { return (k * U) / G; }
__host__ __device__ __forceinline__ int sk_owner(long long u, long long U, int G) { return (int)(((u + 1) * G - 1) / U); }

template <int I, int THREADS, int MINB, int UNROLL, int STAGE_BLKS, int FORM, bool GUARD, int DIMS = 3>
__global__ void __launch_bounds__(THREADS, MINB) force_f32_streamk_kernel(const FastArgs a)
{
    constexpr int SRC_ELEMS = BLK_ELEMS;
    using RingT = Ring<SRC_ELEMS, STAGE_BLKS>;
    constexpr int LANES_PER_BLK = BLK / THREADS;
    static_assert(BLK % THREADS == 0 && I % LANES_PER_BLK == 0, "tile shape");
    constexpr int TILE_BLKS = I / LANES_PER_BLK;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int S = (a.j_nblk + STAGE_BLKS - 1) / STAGE_BLKS;              // source stages per tile
    const long long U = (long long)a.n_tiles * S;
    const int G = (int)gridDim.x;
    const long long g0 = sk_start(blockIdx.x, U, G), g1 = sk_start((long long)blockIdx.x + 1, U, G);
    const int nst = (int)(g1 - g0);
    if (nst <= 0) return;                                                // more CTAs than units
    RingT ring;
    ring.setup(smem_raw, THREADS / 32);
    const int tid = threadIdx.x;
    const float *src = a.posm + (size_t)a.j_blk0 * SRC_ELEMS;
    const int jst0 = (int)(g0 % S);
    auto issue = [&](int t) {                                            // stage t of this CTA's run
        const int jst = (jst0 + t) % S;
        const int nb = min(STAGE_BLKS, a.j_nblk - jst * STAGE_BLKS);
        ring.issue_raw(t % NSTAGE, src + (size_t)jst * STAGE_BLKS * SRC_ELEMS, (uint32_t)nb * SRC_ELEMS * 4u);
    };
    if (tid == 0) {
        const int pre = min(NSTAGE, nst);
        for (int t = 0; t < pre; ++t) issue(t);
    }

    // second-level accumulators (blocked summation): every SK_FLUSH stages the register sums are folded into per-thread
    // shared-memory cells and restarted, so no fp32 accumulator ever adds more than SK_FLUSH x 256 terms per packed lane
    // before it is added to a sum of its own size class -- the error of a long sequential fp32 sum grows with its length
    // (at N = 1M one accumulator per target and lane would see 524,288 terms; the split form had 13 x 2 accumulators)
    float *sacc = reinterpret_cast<float *>(smem_raw + SK_SACC_OFFSET(RingT::SMEM));     // [3][I][THREADS]
    const float2 e2 = make_float2(a.eps2, a.eps2);
    int tile = (int)(g0 / S), jst = jst0, t = 0;
    while (t < nst) {
        // ---- one segment: this CTA's share of `tile`, source stages [jst, jst + seg_n)
        const int seg_n = min(nst - t, S - jst);
        float2 nxi[I], nyi[I], nzi[I], ax[I], ay[I], az[I];
#pragma unroll
        for (int k = 0; k < I; ++k) {
            const int blk = tile * TILE_BLKS + k / LANES_PER_BLK;
            const int lane = (k % LANES_PER_BLK) * THREADS + tid;
            const float *b = a.posm + (size_t)a.i_blk0 * BLK_ELEMS + (size_t)blk * BLK_ELEMS + lane;
            const float x = b[0], y = b[BLK], z = b[2 * BLK];
            nxi[k] = make_float2(-x, -x);
            nyi[k] = make_float2(-y, -y);
            nzi[k] = make_float2(-z, -z);
            ax[k] = ay[k] = az[k] = make_float2(0.f, 0.f);
            sacc[(0 * I + k) * THREADS + tid] = 0.f; sacc[(1 * I + k) * THREADS + tid] = 0.f; sacc[(2 * I + k) * THREADS + tid] = 0.f;
        }
        for (int u = 0; u < seg_n; ++u, ++t) {
            const int s = t % NSTAGE;
            mbar_wait(&ring.full[s], (uint32_t)(t / NSTAGE) & 1u);
            const float *st = ring.stage + (size_t)s * RingT::STAGE_FLOATS;
            const int nb = min(STAGE_BLKS, a.j_nblk - (jst + u) * STAGE_BLKS);
            for (int b = 0; b < nb; ++b) {
                const float *sx = st + b * SRC_ELEMS;
#pragma unroll UNROLL
                for (int j = 0; j < BLK; j += 4) {
                    const float4 X = *reinterpret_cast<const float4 *>(sx + j);
                    const float4 Y = *reinterpret_cast<const float4 *>(sx + BLK + j);
                    const float4 Z = (DIMS == 3) ? *reinterpret_cast<const float4 *>(sx + 2 * BLK + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                    if (FORM == FORM_UNIFORM) {
                        interact_pair_uniform<I, GUARD, DIMS>(lo2(X), lo2(Y), lo2(Z), nxi, nyi, nzi, ax, ay, az, e2);
                        interact_pair_uniform<I, GUARD, DIMS>(hi2(X), hi2(Y), hi2(Z), nxi, nyi, nzi, ax, ay, az, e2);
                        continue;
                    }
                    const float4 M = *reinterpret_cast<const float4 *>(sx + 3 * BLK + j);
                    interact_pair_plain<I, GUARD, DIMS>(lo2(X), lo2(Y), lo2(Z), lo2(M), nxi, nyi, nzi, ax, ay, az, e2);
                    interact_pair_plain<I, GUARD, DIMS>(hi2(X), hi2(Y), hi2(Z), hi2(M), nxi, nyi, nzi, ax, ay, az, e2);
                }
            }
            // this stage's buffer is free; thread 0 refills the buffer of the PREVIOUS stage (see Ring::release_and_refill)
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(&ring.empty[s]);
            if (tid == 0 && t >= 1 && (t - 1 + NSTAGE) < nst) {
                const int tp = t - 1;
                mbar_wait(&ring.empty[tp % NSTAGE], (uint32_t)(tp / NSTAGE) & 1u);
                issue(tp + NSTAGE);
            }
            if (((u + 1) % SK_FLUSH) == 0 && u + 1 < seg_n) {
#pragma unroll
                for (int k = 0; k < I; ++k) {
                    sacc[(0 * I + k) * THREADS + tid] += ax[k].x + ax[k].y;
                    sacc[(1 * I + k) * THREADS + tid] += ay[k].x + ay[k].y;
                    sacc[(2 * I + k) * THREADS + tid] += az[k].x + az[k].y;
                    ax[k] = ay[k] = az[k] = make_float2(0.f, 0.f);
                }
            }
        }
        // ---- the segment's partial sums -> the slot of this CTA among those that share the tile
        const int slot = a.slot0 + (int)blockIdx.x - sk_owner((long long)tile * S, U, G);
#pragma unroll
        for (int k = 0; k < I; ++k) {
            const float fx = sacc[(0 * I + k) * THREADS + tid] + (ax[k].x + ax[k].y);
            const float fy = sacc[(1 * I + k) * THREADS + tid] + (ay[k].x + ay[k].y);
            const float fz = sacc[(2 * I + k) * THREADS + tid] + (az[k].x + az[k].y);
            const int blk = tile * TILE_BLKS + k / LANES_PER_BLK;
            const int lane = (k % LANES_PER_BLK) * THREADS + tid;
            float *o = a.accp + (size_t)slot * a.n_iblk_shard * BLK_ELEMS + (size_t)(a.i_blk_local0 + blk) * BLK_ELEMS + lane;
            o[0] = fx; o[BLK] = fy; o[2 * BLK] = fz;
        }
        jst += seg_n;
        if (jst == S) { jst = 0; ++tile; }
    }
}

} // namespace nb
