// force_f32.cu -- all-pairs softened-gravity force accumulation, fp32, hand-written for sm_100a.
//
// Restates the reference's per-pair kernel, Quadtree::acc leaf loop (Quadtree.hpp:133-144) with
// Quadtree::fast_inv_sqrt (Quadtree.hpp:106-111), driven over all targets as Simulation::attract
// does (Simulation.hpp:203-207).  Two variants:
//
//  * force_f32_fast_kernel      headline path.  256 threads, 4 target bodies per thread held in
//    registers; source tiles (blocked SoA, 8 KiB per stage) are staged into shared memory by 1-D
//    TMA bulk copies (cp.async.bulk + mbarrier full/empty ring, 4 stages); the inner loop runs on
//    Blackwell's packed-fp32 instructions (FADD2/FFMA2/FMUL2: two sources per issue slot) with
//    one MUFU.RSQ per interaction -- 12 fp32-pipe lane-ops + 1 MUFU per interaction and ~0.6 issue
//    slots per lane-op, so the FP32 datapath, not the issue port, is the limiter.  No tensor
//    cores: this is not a dense contraction.
//  * force_f32_refcompat_kernel parity path.  Bit-faithful restatement: unfused IEEE mul/add in
//    the reference's expression order, the 0x5f3759df bit trick + one Newton step, the r_sq > 0
//    guard, and accumulation over sources in index order 0..n-1 by a single thread per target,
//    so accelerations equal the reference's (strict build) bit for bit.
#include "kernels.h"

namespace nb {

constexpr int STAGE_BLKS = 2;                       // source blocks per pipeline stage (512 bodies)
constexpr int NSTAGE = 4;                           // ring depth
constexpr int STAGE_FLOATS = STAGE_BLKS * BLK_ELEMS;
constexpr int STAGE_BYTES = STAGE_FLOATS * 4;       // 8 KiB
constexpr size_t F32_SMEM = (size_t)NSTAGE * STAGE_BYTES + 2 * NSTAGE * sizeof(uint64_t);

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
struct Ring {
    float *stage;        // NSTAGE * STAGE_FLOATS
    uint64_t *full;      // NSTAGE, tx-count barriers armed by the producer thread
    uint64_t *empty;     // NSTAGE, one arrival per consumer warp
};

__device__ __forceinline__ void ring_issue(const Ring &r, const float *src_blocks, int t,
                                           int chunk_blks)
{
    const int s = t % NSTAGE;
    const int nb = min(STAGE_BLKS, chunk_blks - t * STAGE_BLKS);
    const uint32_t bytes = (uint32_t)nb * BLK_ELEMS * 4u;
    mbar_expect_tx(&r.full[s], bytes);
    tma_bulk_g2s(r.stage + (size_t)s * STAGE_FLOATS,
                 src_blocks + (size_t)t * STAGE_BLKS * BLK_ELEMS, bytes, &r.full[s]);
}

__device__ __forceinline__ Ring ring_setup(unsigned char *smem_raw, int nwarps)
{
    Ring r;
    r.stage = reinterpret_cast<float *>(smem_raw);
    r.full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)NSTAGE * STAGE_BYTES);
    r.empty = r.full + NSTAGE;
    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; ++s) {
            mbar_init(&r.full[s], 1);
            mbar_init(&r.empty[s], nwarps);
        }
        mbar_fence_init();
    }
    __syncthreads();
    return r;
}

// ---- fast kernel --------------------------------------------------------------------------------
// 4 sources (two f32x2 pairs) against this thread's I targets.
template <int I, bool GUARD>
__device__ __forceinline__ void interact4(const float4 X, const float4 Y, const float4 Z,
                                          const float4 M, const float2 (&nxi)[I],
                                          const float2 (&nyi)[I], const float2 (&nzi)[I],
                                          float2 (&ax)[I], float2 (&ay)[I], float2 (&az)[I],
                                          const float2 e2)
{
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const float2 xj = h ? make_float2(X.z, X.w) : make_float2(X.x, X.y);
        const float2 yj = h ? make_float2(Y.z, Y.w) : make_float2(Y.x, Y.y);
        const float2 zj = h ? make_float2(Z.z, Z.w) : make_float2(Z.x, Z.y);
        const float2 mj = h ? make_float2(M.z, M.w) : make_float2(M.x, M.y);
#pragma unroll
        for (int k = 0; k < I; ++k) {
            const float2 dx = __fadd2_rn(xj, nxi[k]);            // r = p_j - p_i   (FADD2)
            const float2 dy = __fadd2_rn(yj, nyi[k]);
            const float2 dz = __fadd2_rn(zj, nzi[k]);
            float2 r2 = __ffma2_rn(dx, dx, e2);                  // r^2 + eps^2     (3 FFMA2)
            r2 = __ffma2_rn(dy, dy, r2);
            r2 = __ffma2_rn(dz, dz, r2);
            float2 ri = make_float2(rsqrt_approx(r2.x), rsqrt_approx(r2.y)); // 2 MUFU.RSQ
            if (GUARD) { // eps == 0: self / coincident pairs contribute nothing (Quadtree.hpp:139)
                ri.x = (r2.x > 0.0f) ? ri.x : 0.0f;
                ri.y = (r2.y > 0.0f) ? ri.y : 0.0f;
            }
            const float2 ri2 = __fmul2_rn(ri, ri);
            const float2 mr = __fmul2_rn(mj, ri);
            const float2 s = __fmul2_rn(mr, ri2);                // m / (r^2+eps^2)^(3/2)
            ax[k] = __ffma2_rn(dx, s, ax[k]);                    // acc += r * s    (3 FFMA2)
            ay[k] = __ffma2_rn(dy, s, ay[k]);
            az[k] = __ffma2_rn(dz, s, az[k]);
        }
    }
}

template <bool GUARD, bool FUSE>
__global__ void __launch_bounds__(FAST_THREADS, 2)
force_f32_fast_kernel(const float *__restrict__ posm, float *__restrict__ accp, int i_blk0,
                      int i_blk_local0, int n_iblk_shard, int j_blk0, int j_nblk, int splits,
                      int slot0, float eps2, long long n_real, float *__restrict__ posm_next,
                      float *__restrict__ vel, float *__restrict__ acc, IntegParams ip)
{
    constexpr int I = FAST_I;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const Ring ring = ring_setup(smem_raw, FAST_THREADS / 32);

    const int tile = blockIdx.x / splits;
    const int split = blockIdx.x - tile * splits;
    const int tid = threadIdx.x;

    // source chunk of this CTA: whole blocks, balanced to within one block
    const int jb0 = j_blk0 + (int)(((long long)j_nblk * split) / splits);
    const int jb1 = j_blk0 + (int)(((long long)j_nblk * (split + 1)) / splits);
    const int chunk_blks = jb1 - jb0;
    const int nst = (chunk_blks + STAGE_BLKS - 1) / STAGE_BLKS;
    const float *src = posm + (size_t)jb0 * BLK_ELEMS;

    if (tid == 0) {
        const int pre = min(NSTAGE, nst);
        for (int t = 0; t < pre; ++t) ring_issue(ring, src, t, chunk_blks);
    }

    // targets: body (tile*I + k)*256 + tid of the launch, k = 0..I-1 -> coalesced block reads.
    // Keep -p_i broadcast over both packed lanes so that r = p_j + (-p_i) is one FADD2.
    float2 nxi[I], nyi[I], nzi[I], ax[I], ay[I], az[I];
#pragma unroll
    for (int k = 0; k < I; ++k) {
        const float *b = posm + (size_t)(i_blk0 + tile * I + k) * BLK_ELEMS + tid;
        const float x = b[0], y = b[BLK], z = b[2 * BLK];
        nxi[k] = make_float2(-x, -x);
        nyi[k] = make_float2(-y, -y);
        nzi[k] = make_float2(-z, -z);
        ax[k] = ay[k] = az[k] = make_float2(0.f, 0.f);
    }
    const float2 e2 = make_float2(eps2, eps2);

    for (int t = 0; t < nst; ++t) {
        const int s = t % NSTAGE;
        mbar_wait(&ring.full[s], (uint32_t)(t / NSTAGE) & 1u);
        const float *st = ring.stage + (size_t)s * STAGE_FLOATS;
        const int nb = min(STAGE_BLKS, chunk_blks - t * STAGE_BLKS);
        for (int b = 0; b < nb; ++b) {
            const float *sx = st + b * BLK_ELEMS;
#pragma unroll 2
            for (int j = 0; j < BLK; j += 4) {
                const float4 X = *reinterpret_cast<const float4 *>(sx + j);
                const float4 Y = *reinterpret_cast<const float4 *>(sx + BLK + j);
                const float4 Z = *reinterpret_cast<const float4 *>(sx + 2 * BLK + j);
                const float4 M = *reinterpret_cast<const float4 *>(sx + 3 * BLK + j);
                interact4<I, GUARD>(X, Y, Z, M, nxi, nyi, nzi, ax, ay, az, e2);
            }
        }
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&ring.empty[s]);
        // refill the buffer of the PREVIOUS stage (all warps have almost surely left it by now)
        if (tid == 0 && t >= 1 && (t - 1 + NSTAGE) < nst) {
            const int tp = t - 1;
            mbar_wait(&ring.empty[tp % NSTAGE], (uint32_t)(tp / NSTAGE) & 1u);
            ring_issue(ring, src, tp + NSTAGE, chunk_blks);
        }
    }

    // epilogue: fold the two packed lanes (even/odd sources)
#pragma unroll
    for (int k = 0; k < I; ++k) {
        const float fx = ax[k].x + ax[k].y, fy = ay[k].x + ay[k].y, fz = az[k].x + az[k].y;
        const int lb = i_blk_local0 + tile * I + k; // block inside the shard
        if (FUSE) {
            if ((long long)(i_blk0 + tile * I + k) * BLK + tid >= n_real) continue; // padding stays put
            // fused kick-drift: the new positions go to the other posm buffer, so CTAs still
            // reading the current one are undisturbed (race-free by construction).
            const float gx = fx * ip.G, gy = fy * ip.G, gz = fz * ip.G;
            const float *pb = posm + (size_t)(i_blk0 + tile * I + k) * BLK_ELEMS + tid;
            float *vb = vel + (size_t)lb * BLK_ELEMS + tid;
            float *ab = acc + (size_t)lb * BLK_ELEMS + tid;
            float *nb_ = posm_next + (size_t)(i_blk0 + tile * I + k) * BLK_ELEMS + tid;
            float px = pb[0], py = pb[BLK], pz = pb[2 * BLK];
            const float m = pb[3 * BLK];
            float vx = vb[0], vy = vb[BLK], vz = vb[2 * BLK];
            integrate_body_f32(px, py, pz, vx, vy, vz, gx, gy, gz, ip);
            nb_[0] = px; nb_[BLK] = py; nb_[2 * BLK] = pz; nb_[3 * BLK] = m;
            vb[0] = vx; vb[BLK] = vy; vb[2 * BLK] = vz;
            ab[0] = gx; ab[BLK] = gy; ab[2 * BLK] = gz;
        } else {
            float *o = accp + ((size_t)(slot0 + split) * n_iblk_shard + lb) * BLK_ELEMS + tid;
            o[0] = fx; o[BLK] = fy; o[2 * BLK] = fz;
        }
    }
}

// ---- refcompat kernel ---------------------------------------------------------------------------
// Quadtree.hpp:106-111, every operation individually rounded (no FMA contraction).
__device__ __forceinline__ float quake_inv_sqrt(float number)
{
    const float y = __uint_as_float(0x5f3759dfu - (__float_as_uint(number) >> 1));
    const float t = __fmul_rn(__fmul_rn(__fmul_rn(number, 0.5f), y), y);
    return __fmul_rn(y, __fsub_rn(1.5f, t));
}

__device__ __forceinline__ void ref_pair(float xj, float yj, float zj, float mj, float xi, float yi,
                                         float zi, float e_sq, float &ax, float &ay, float &az)
{
    const float rx = __fsub_rn(xj, xi), ry = __fsub_rn(yj, yi), rz = __fsub_rn(zj, zi);
    // Vec2::mag_sq = x*x + y*y (Vec2.hpp:216-219); the z term is appended, and is +0 in 2-D
    const float r_sq =
        __fadd_rn(__fadd_rn(__fmul_rn(rx, rx), __fmul_rn(ry, ry)), __fmul_rn(rz, rz));
    if (r_sq > 0.0f) {
        const float inv = quake_inv_sqrt(__fadd_rn(r_sq, e_sq));
        const float inv3 = __fmul_rn(__fmul_rn(inv, inv), inv);
        const float s = __fmul_rn(mj, inv3);
        ax = __fadd_rn(ax, __fmul_rn(rx, s));
        ay = __fadd_rn(ay, __fmul_rn(ry, s));
        az = __fadd_rn(az, __fmul_rn(rz, s));
    }
}

__global__ void __launch_bounds__(REF_THREADS, 4)
force_f32_refcompat_kernel(const float *__restrict__ posm, float *__restrict__ accp, int i_blk0,
                           int i_blk_local0, int n_iblk_shard, int j_blk0, int j_nblk,
                           long long j_body_limit, int slot0, float eps2)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const Ring ring = ring_setup(smem_raw, REF_THREADS / 32);
    const int tid = threadIdx.x;
    const int chunk_blks = j_nblk;
    const int nst = (chunk_blks + STAGE_BLKS - 1) / STAGE_BLKS;
    const float *src = posm + (size_t)j_blk0 * BLK_ELEMS;
    if (tid == 0) {
        const int pre = min(NSTAGE, nst);
        for (int t = 0; t < pre; ++t) ring_issue(ring, src, t, chunk_blks);
    }
    // 128 targets per CTA: half a block
    const int half = blockIdx.x;                      // half-block index inside the launch
    const size_t ib = (size_t)i_blk0 + (half >> 1);
    const int lane = (half & 1) * REF_TILE_BODIES + tid;
    const float *b = posm + ib * BLK_ELEMS + lane;
    const float xi = b[0], yi = b[BLK], zi = b[2 * BLK];
    float ax = 0.f, ay = 0.f, az = 0.f;

    for (int t = 0; t < nst; ++t) {
        const int s = t % NSTAGE;
        mbar_wait(&ring.full[s], (uint32_t)(t / NSTAGE) & 1u);
        const float *st = ring.stage + (size_t)s * STAGE_FLOATS;
        const int nb = min(STAGE_BLKS, chunk_blks - t * STAGE_BLKS);
        for (int bb = 0; bb < nb; ++bb) {
            const float *sx = st + bb * BLK_ELEMS;
            const long long first = (long long)(j_blk0 + t * STAGE_BLKS + bb) * BLK;
            const long long rem = j_body_limit - first;   // real (non-padding) sources in block
            const int cnt = rem >= BLK ? BLK : (rem > 0 ? (int)rem : 0);
            int j = 0;
            for (; j + 4 <= cnt; j += 4) { // sources strictly in index order
                const float4 X = *reinterpret_cast<const float4 *>(sx + j);
                const float4 Y = *reinterpret_cast<const float4 *>(sx + BLK + j);
                const float4 Z = *reinterpret_cast<const float4 *>(sx + 2 * BLK + j);
                const float4 M = *reinterpret_cast<const float4 *>(sx + 3 * BLK + j);
                ref_pair(X.x, Y.x, Z.x, M.x, xi, yi, zi, eps2, ax, ay, az);
                ref_pair(X.y, Y.y, Z.y, M.y, xi, yi, zi, eps2, ax, ay, az);
                ref_pair(X.z, Y.z, Z.z, M.z, xi, yi, zi, eps2, ax, ay, az);
                ref_pair(X.w, Y.w, Z.w, M.w, xi, yi, zi, eps2, ax, ay, az);
            }
            for (; j < cnt; ++j)
                ref_pair(sx[j], sx[BLK + j], sx[2 * BLK + j], sx[3 * BLK + j], xi, yi, zi, eps2, ax,
                         ay, az);
        }
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&ring.empty[s]);
        if (tid == 0 && t >= 1 && (t - 1 + NSTAGE) < nst) {
            const int tp = t - 1;
            mbar_wait(&ring.empty[tp % NSTAGE], (uint32_t)(tp / NSTAGE) & 1u);
            ring_issue(ring, src, tp + NSTAGE, chunk_blks);
        }
    }
    const int lb = i_blk_local0 + (half >> 1);
    float *o = accp + ((size_t)slot0 * n_iblk_shard + lb) * BLK_ELEMS + lane;
    o[0] = ax; o[BLK] = ay; o[2 * BLK] = az;
}

// ---- host-side launchers ------------------------------------------------------------------------
template <bool GUARD, bool FUSE>
static cudaError_t launch_fast_t(const ForceLaunch &L, cudaStream_t st)
{
    auto kern = force_f32_fast_kernel<GUARD, FUSE>;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)F32_SMEM);
        if (e != cudaSuccess) return e;
        attr_done = true;
    }
    const int grid = force_f32_fast_grid(L);
    kern<<<grid, FAST_THREADS, F32_SMEM, st>>>(
        (const float *)L.posm, (float *)L.accp, L.i_blk0, L.i_blk_local0, L.n_iblk_shard, L.j_blk0,
        L.j_nblk, L.splits, L.slot0, L.eps2, L.j_body_limit, (float *)L.posm_next, (float *)L.vel,
        (float *)L.acc, L.ip);
    return cudaGetLastError();
}

int force_f32_fast_grid(const ForceLaunch &L) { return (L.n_iblk / FAST_TILE_BLKS) * L.splits; }

cudaError_t launch_force_f32_fast(const ForceLaunch &L, bool guard_zero, cudaStream_t st)
{
    if (L.n_iblk % FAST_TILE_BLKS != 0 || L.splits < 1 || L.j_nblk < L.splits)
        return cudaErrorInvalidValue;
    if (L.fuse && L.splits != 1) return cudaErrorInvalidValue;
    if (guard_zero) return L.fuse ? launch_fast_t<true, true>(L, st) : launch_fast_t<true, false>(L, st);
    return L.fuse ? launch_fast_t<false, true>(L, st) : launch_fast_t<false, false>(L, st);
}

int force_f32_fast_ctas_per_sm(bool fuse)
{
    int n = 0;
    cudaError_t e;
    if (fuse)
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, force_f32_fast_kernel<false, true>,
                                                          FAST_THREADS, F32_SMEM);
    else
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, force_f32_fast_kernel<false, false>,
                                                          FAST_THREADS, F32_SMEM);
    return e == cudaSuccess ? n : 0;
}

cudaError_t launch_force_f32_refcompat(const ForceLaunch &L, cudaStream_t st)
{
    if (L.splits != 1 || L.j_nblk < 1) return cudaErrorInvalidValue;
    const int grid = L.n_iblk * (BLK / REF_TILE_BODIES);
    force_f32_refcompat_kernel<<<grid, REF_THREADS, F32_SMEM, st>>>(
        (const float *)L.posm, (float *)L.accp, L.i_blk0, L.i_blk_local0, L.n_iblk_shard, L.j_blk0,
        L.j_nblk, L.j_body_limit, L.slot0, L.eps2);
    return cudaGetLastError();
}

// ---- stand-alone integrator (vectorised, coalesced) ---------------------------------------------
// One thread per 4 consecutive bodies of a block: float4 loads/stores on every component array.
// acc = G * sum over partial slots in slot order; then the same integrate_body_f32 as the fused
// epilogue.  Bytes per body: read posm 16 + vel 12 + 12*nslots, write posm 16 + vel 12 + acc 12.
__global__ void __launch_bounds__(256)
integrate_f32_kernel(const float *__restrict__ posm_cur, float *__restrict__ posm_next,
                     float *__restrict__ vel, float *__restrict__ acc,
                     const float *__restrict__ accp, int nslots, int i_blk0, int n_iblk_shard,
                     int acc_only, long long n_real, IntegParams ip)
{
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;   // one per 4 bodies
    const int lb = gid >> 6;                                 // 64 float4 groups per block
    if (lb >= n_iblk_shard) return;
    const int q = (gid & 63) * 4;
    const size_t loff = (size_t)lb * BLK_ELEMS + q;
    const size_t goff = (size_t)(i_blk0 + lb) * BLK_ELEMS + q;

    float4 A[3] = {make_float4(0, 0, 0, 0), make_float4(0, 0, 0, 0), make_float4(0, 0, 0, 0)};
    for (int s = 0; s < nslots; ++s) {
        const float *p = accp + (size_t)s * n_iblk_shard * BLK_ELEMS + loff;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float4 v = *reinterpret_cast<const float4 *>(p + c * BLK);
            A[c].x += v.x; A[c].y += v.y; A[c].z += v.z; A[c].w += v.w;
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        A[c].x *= ip.G; A[c].y *= ip.G; A[c].z *= ip.G; A[c].w *= ip.G;
        *reinterpret_cast<float4 *>(acc + loff + c * BLK) = A[c];
    }
    if (acc_only) return;
    const long long body0 = (long long)(i_blk0 + lb) * BLK + q;
    if (body0 >= n_real) return;                             // pure padding: never moved

    float4 P[4], V[3];
#pragma unroll
    for (int c = 0; c < 4; ++c) P[c] = *reinterpret_cast<const float4 *>(posm_cur + goff + c * BLK);
#pragma unroll
    for (int c = 0; c < 3; ++c) V[c] = *reinterpret_cast<const float4 *>(vel + loff + c * BLK);
    integrate_body_f32(P[0].x, P[1].x, P[2].x, V[0].x, V[1].x, V[2].x, A[0].x, A[1].x, A[2].x, ip);
    if (body0 + 1 < n_real)
        integrate_body_f32(P[0].y, P[1].y, P[2].y, V[0].y, V[1].y, V[2].y, A[0].y, A[1].y, A[2].y, ip);
    if (body0 + 2 < n_real)
        integrate_body_f32(P[0].z, P[1].z, P[2].z, V[0].z, V[1].z, V[2].z, A[0].z, A[1].z, A[2].z, ip);
    if (body0 + 3 < n_real)
        integrate_body_f32(P[0].w, P[1].w, P[2].w, V[0].w, V[1].w, V[2].w, A[0].w, A[1].w, A[2].w, ip);
#pragma unroll
    for (int c = 0; c < 4; ++c) *reinterpret_cast<float4 *>(posm_next + goff + c * BLK) = P[c];
#pragma unroll
    for (int c = 0; c < 3; ++c) *reinterpret_cast<float4 *>(vel + loff + c * BLK) = V[c];
}

cudaError_t launch_integrate_f32(const IntegLaunch &L, cudaStream_t st)
{
    const int threads = L.n_iblk_shard * 64;
    const int grid = (threads + 255) / 256;
    integrate_f32_kernel<<<grid, 256, 0, st>>>((const float *)L.posm_cur, (float *)L.posm_next,
                                               (float *)L.vel, (float *)L.acc,
                                               (const float *)L.accp, L.nslots, L.i_blk0,
                                               L.n_iblk_shard, L.acc_only, L.n_real, L.ip);
    return cudaGetLastError();
}

} // namespace nb
