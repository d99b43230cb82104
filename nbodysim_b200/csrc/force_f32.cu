// force_f32.cu -- all-pairs softened-gravity force accumulation, fp32, hand-written for sm_100a.
//
// Restates the reference's per-pair kernel, Quadtree::acc leaf loop (Quadtree.hpp:133-144) with
// Quadtree::fast_inv_sqrt (Quadtree.hpp:106-111), driven over all targets as Simulation::attract
// does (Simulation.hpp:203-207).  Two variants:
//
//  * force_f32_fast_kernel (force_f32_fast.cuh)  headline path.  256 threads, 8 target bodies per
//    thread held in registers, one CTA per SM; source tiles (blocked SoA, 8 KiB per stage) are
//    staged into shared memory by 1-D TMA bulk copies (cp.async.bulk + mbarrier full/empty ring,
//    4 stages); the inner loop runs on Blackwell's packed-fp32 instructions (FADD2/FFMA2/FMUL2:
//    two sources per issue slot) with one MUFU.RSQ per interaction: 12 fp32-pipe lane-ops per
//    interaction in general, 11 when all massive bodies share one mass (the multiply by m_j then
//    factors out of the sum -- the "uniform" form, chosen automatically at upload).  Measured on
//    B200 (tools/ubench.cu, tools/kbench.cu): an FFMA2 occupies the FP32 pipe for 2 cycles (same
//    datapath peak as FFMA, half the issue slots), an FFMA2 with three distinct register pairs
//    costs ~2.6 cycles (register-file read bandwidth) and each MUFU.RSQ steals ~1-2 cycles, which
//    bounds these instruction mixes at ~73 % / ~79 % of the 20-flop FP32 peak.
//    No tensor cores: this is not a dense contraction.
//  * force_f32_refcompat_kernel  parity path.  Bit-faithful restatement: unfused IEEE mul/add in
//    the reference's expression order, the 0x5f3759df bit trick + one Newton step, the r_sq > 0
//    guard, and accumulation over sources in index order 0..n-1 by a single thread per target,
//    so accelerations equal the reference's (strict build) bit for bit.
#include "force_f32_fast.cuh"
#include <cstring>

namespace nb {

using RefRing = Ring<BLK_ELEMS, 2>;
constexpr int STAGE_BLKS = 2;

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
    RefRing ring;
    ring.setup(smem_raw, REF_THREADS / 32);
    const int tid = threadIdx.x;
    const int chunk_blks = j_nblk;
    const int nst = (chunk_blks + STAGE_BLKS - 1) / STAGE_BLKS;
    const float *src = posm + (size_t)j_blk0 * BLK_ELEMS;
    if (tid == 0) {
        const int pre = min(NSTAGE, nst);
        for (int t = 0; t < pre; ++t) ring.issue(src, t, chunk_blks);
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
        const float *st = ring.stage + (size_t)s * RefRing::STAGE_FLOATS;
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
        ring.release_and_refill(src, t, nst, chunk_blks);
    }
    const int lb = i_blk_local0 + (half >> 1);
    float *o = accp + ((size_t)slot0 * n_iblk_shard + lb) * BLK_ELEMS + lane;
    o[0] = ax; o[BLK] = ay; o[2 * BLK] = az;
}

// ---- host-side launchers ------------------------------------------------------------------------
struct LargeCfg {
    static constexpr int I = FAST_I, THREADS = FAST_THREADS, MINB = FAST_MINB, UNROLL = FAST_UNROLL,
                         STAGE = FAST_STAGE_BLKS, TILE_BLKS = FAST_TILE_BLKS;
};
struct SmallCfg {
    static constexpr int I = SMALL_I, THREADS = SMALL_THREADS, MINB = SMALL_MINB, UNROLL = SMALL_UNROLL,
                         STAGE = SMALL_STAGE_BLKS, TILE_BLKS = SMALL_TILE_BLKS;
};

template <typename C, int FORM, bool GUARD, bool FUSE, int DIMS>
static cudaError_t launch_fast_d(const ForceLaunch &L, cudaStream_t st)
{
    using RingT = Ring<BLK_ELEMS, C::STAGE>;
    auto kern = force_f32_fast_kernel<C::I, C::THREADS, C::MINB, C::UNROLL, C::STAGE, FORM, GUARD, FUSE, DIMS>;
    static_assert(RingT::SMEM <= 48 * 1024, "above 48 KiB the per-device opt-in attribute would be needed");
    FastArgs a;
    a.posm = (const float *)L.posm;
    a.accp = (float *)L.accp;
    a.i_blk0 = L.i_blk0; a.i_blk_local0 = L.i_blk_local0; a.n_iblk_shard = L.n_iblk_shard;
    a.j_blk0 = L.j_blk0; a.j_nblk = L.j_nblk; a.splits = L.splits; a.slot0 = L.slot0;
    a.eps2 = L.eps2; a.acc_scale = L.acc_scale; a.n_real = L.j_body_limit;
    a.posm_next = (float *)L.posm_next; a.vel = (float *)L.vel; a.acc = (float *)L.acc; a.ip = L.ip;
    kern<<<(L.n_iblk / C::TILE_BLKS) * L.splits, C::THREADS, RingT::SMEM, st>>>(a);
    return cudaGetLastError();
}

template <typename C, int FORM, bool GUARD, int DIMS>
static cudaError_t launch_streamk_d(const ForceLaunch &L, cudaStream_t st)
{
    using RingT = Ring<BLK_ELEMS, C::STAGE>;
    auto kern = force_f32_streamk_kernel<C::I, C::THREADS, C::MINB, C::UNROLL, C::STAGE, FORM, GUARD, DIMS>;
    FastArgs a;
    memset(&a, 0, sizeof a);
    a.posm = (const float *)L.posm;
    a.accp = (float *)L.accp;
    a.i_blk0 = L.i_blk0; a.i_blk_local0 = L.i_blk_local0; a.n_iblk_shard = L.n_iblk_shard;
    a.j_blk0 = L.j_blk0; a.j_nblk = L.j_nblk; a.splits = 1; a.slot0 = L.slot0;
    a.n_tiles = L.n_iblk / C::TILE_BLKS;
    a.eps2 = L.eps2; a.acc_scale = L.acc_scale; a.n_real = L.j_body_limit;
    constexpr size_t smem = SK_SMEM(RingT::SMEM, C::I, C::THREADS);
    if (smem > 48 * 1024) {                       // the ring plus the second-level accumulators: opt in once per kernel
        static const cudaError_t attr = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (attr != cudaSuccess) return attr;
    }
    kern<<<L.streamk_ctas, C::THREADS, smem, st>>>(a);
    return cudaGetLastError();
}
template <int FORM, bool GUARD>
static cudaError_t launch_streamk_t(const ForceLaunch &L, cudaStream_t st)
{
    if (L.small_tile) return L.dims == 2 ? launch_streamk_d<SmallCfg, FORM, GUARD, 2>(L, st) : launch_streamk_d<SmallCfg, FORM, GUARD, 3>(L, st);
    return L.dims == 2 ? launch_streamk_d<LargeCfg, FORM, GUARD, 2>(L, st) : launch_streamk_d<LargeCfg, FORM, GUARD, 3>(L, st);
}

int force_f32_streamk_owner(long long unit, long long units, int ctas) { return sk_owner(unit, units, ctas); }

int force_f32_streamk_slots(int tiles, int stages, int ctas)
{
    const long long U = (long long)tiles * stages;
    int worst = 1;
    for (int t = 0; t < tiles; ++t)
        worst = std::max(worst, sk_owner((long long)t * stages + stages - 1, U, ctas) - sk_owner((long long)t * stages, U, ctas) + 1);
    return worst;
}

int force_f32_streamk_ctas_per_sm(bool uniform_mass, bool small_tile)
{
    int n = 0;
    cudaError_t e;
    if (small_tile) {
        using RingT = Ring<BLK_ELEMS, SMALL_STAGE_BLKS>;
        e = uniform_mass ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, force_f32_streamk_kernel<SMALL_I, SMALL_THREADS, SMALL_MINB, SMALL_UNROLL, SMALL_STAGE_BLKS, FORM_UNIFORM, false, 3>, SMALL_THREADS, SK_SMEM(RingT::SMEM, SMALL_I, SMALL_THREADS))
                         : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, force_f32_streamk_kernel<SMALL_I, SMALL_THREADS, SMALL_MINB, SMALL_UNROLL, SMALL_STAGE_BLKS, FORM_PLAIN, false, 3>, SMALL_THREADS, SK_SMEM(RingT::SMEM, SMALL_I, SMALL_THREADS));
    } else {
        using RingT = Ring<BLK_ELEMS, FAST_STAGE_BLKS>;
        e = uniform_mass ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, force_f32_streamk_kernel<FAST_I, FAST_THREADS, FAST_MINB, FAST_UNROLL, FAST_STAGE_BLKS, FORM_UNIFORM, false, 3>, FAST_THREADS, SK_SMEM(RingT::SMEM, FAST_I, FAST_THREADS))
                         : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, force_f32_streamk_kernel<FAST_I, FAST_THREADS, FAST_MINB, FAST_UNROLL, FAST_STAGE_BLKS, FORM_PLAIN, false, 3>, FAST_THREADS, SK_SMEM(RingT::SMEM, FAST_I, FAST_THREADS));
    }
    return e == cudaSuccess ? n : 0;
}

template <int FORM, bool GUARD, bool FUSE>
static cudaError_t launch_fast_t(const ForceLaunch &L, cudaStream_t st)
{
    if (L.small_tile) { // small tiles never fuse (FUSE instantiations exist for the large tile only)
        return L.dims == 2 ? launch_fast_d<SmallCfg, FORM, GUARD, false, 2>(L, st)
                           : launch_fast_d<SmallCfg, FORM, GUARD, false, 3>(L, st);
    }
    return L.dims == 2 ? launch_fast_d<LargeCfg, FORM, GUARD, FUSE, 2>(L, st)
                       : launch_fast_d<LargeCfg, FORM, GUARD, FUSE, 3>(L, st);
}

int force_f32_fast_grid(const ForceLaunch &L)
{
    return (L.n_iblk / (L.small_tile ? SMALL_TILE_BLKS : FAST_TILE_BLKS)) * L.splits;
}

cudaError_t launch_force_f32_fast(const ForceLaunch &L, bool guard_zero, cudaStream_t st)
{
    const int tile = L.small_tile ? SMALL_TILE_BLKS : FAST_TILE_BLKS;
    if (L.streamk_ctas > 0) {
        if (L.n_iblk % tile != 0 || L.j_nblk < 1 || L.fuse) return cudaErrorInvalidValue;
        switch ((L.uniform_mass ? 2 : 0) | (guard_zero ? 1 : 0)) {
        case 0: return launch_streamk_t<FORM_PLAIN, false>(L, st);
        case 1: return launch_streamk_t<FORM_PLAIN, true>(L, st);
        case 2: return launch_streamk_t<FORM_UNIFORM, false>(L, st);
        default: return launch_streamk_t<FORM_UNIFORM, true>(L, st);
        }
    }
    if (L.n_iblk % tile != 0 || L.splits < 1 || L.j_nblk < L.splits) return cudaErrorInvalidValue;
    if (L.fuse && (L.splits != 1 || L.small_tile)) return cudaErrorInvalidValue;
    const int v = (L.uniform_mass ? 4 : 0) | (guard_zero ? 2 : 0) | (L.fuse ? 1 : 0);
    switch (v) {
    case 0: return launch_fast_t<FORM_PLAIN, false, false>(L, st);
    case 1: return launch_fast_t<FORM_PLAIN, false, true>(L, st);
    case 2: return launch_fast_t<FORM_PLAIN, true, false>(L, st);
    case 3: return launch_fast_t<FORM_PLAIN, true, true>(L, st);
    case 4: return launch_fast_t<FORM_UNIFORM, false, false>(L, st);
    case 5: return launch_fast_t<FORM_UNIFORM, false, true>(L, st);
    case 6: return launch_fast_t<FORM_UNIFORM, true, false>(L, st);
    default: return launch_fast_t<FORM_UNIFORM, true, true>(L, st);
    }
}

int force_f32_fast_ctas_per_sm(bool uniform_mass, bool small_tile)
{
    int n = 0;
    cudaError_t e;
    if (small_tile) {
        using RingT = Ring<BLK_ELEMS, SMALL_STAGE_BLKS>;
        if (uniform_mass)
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(
                &n, force_f32_fast_kernel<SMALL_I, SMALL_THREADS, SMALL_MINB, SMALL_UNROLL, SMALL_STAGE_BLKS, FORM_UNIFORM, false, false, 3>,
                SMALL_THREADS, RingT::SMEM);
        else
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(
                &n, force_f32_fast_kernel<SMALL_I, SMALL_THREADS, SMALL_MINB, SMALL_UNROLL, SMALL_STAGE_BLKS, FORM_PLAIN, false, false, 3>,
                SMALL_THREADS, RingT::SMEM);
    } else {
        using RingT = Ring<BLK_ELEMS, FAST_STAGE_BLKS>;
        if (uniform_mass)
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(
                &n, force_f32_fast_kernel<FAST_I, FAST_THREADS, FAST_MINB, FAST_UNROLL, FAST_STAGE_BLKS, FORM_UNIFORM, false, false, 3>,
                FAST_THREADS, RingT::SMEM);
        else
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(
                &n, force_f32_fast_kernel<FAST_I, FAST_THREADS, FAST_MINB, FAST_UNROLL, FAST_STAGE_BLKS, FORM_PLAIN, false, false, 3>,
                FAST_THREADS, RingT::SMEM);
    }
    return e == cudaSuccess ? n : 0;
}

cudaError_t launch_force_f32_refcompat(const ForceLaunch &L, cudaStream_t st)
{
    if (L.splits != 1 || L.j_nblk < 1) return cudaErrorInvalidValue;
    const int grid = L.n_iblk * (BLK / REF_TILE_BODIES);
    force_f32_refcompat_kernel<<<grid, REF_THREADS, RefRing::SMEM, st>>>(
        (const float *)L.posm, (float *)L.accp, L.i_blk0, L.i_blk_local0, L.n_iblk_shard, L.j_blk0,
        L.j_nblk, L.j_body_limit, L.slot0, L.eps2);
    return cudaGetLastError();
}

// ---- stand-alone integrator (vectorised, coalesced) ---------------------------------------------
// One thread per 4 consecutive bodies of a block: float4 loads/stores on every component array.
// acc = G * sum over partial slots in slot order; then the same integrate_body_f32 as the fused
// epilogue.  Bytes per body: read posm 16 + vel 12 + 12*nslots, write posm 16 + vel 12 + acc 12.
__device__ __forceinline__ void
integrate_f32_group(const float *__restrict__ posm_cur, const PeerDests &dests,
                    float *__restrict__ vel, float *__restrict__ acc,
                    const float *__restrict__ accp, float acc_scale, const SlotPlan &slots, int i_blk0,
                    int n_iblk_shard, int acc_only, long long n_real, const IntegParams &ip)
{
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;   // one per 4 bodies
    const int lb = gid >> 6;                                 // 64 float4 groups per block
    if (lb >= n_iblk_shard) return;
    const int q = (gid & 63) * 4;
    const size_t loff = (size_t)lb * BLK_ELEMS + q;
    const size_t goff = (size_t)(i_blk0 + lb) * BLK_ELEMS + q;

    float4 A[3] = {make_float4(0, 0, 0, 0), make_float4(0, 0, 0, 0), make_float4(0, 0, 0, 0)};
    for (int r = 0; r < slots.n; ++r) {                      // the launches of the step in order, each launch's slots in order
        const SlotRange &R = slots.r[r];
        int ns = R.nslots;
        if (R.sk_S > 0) {                                    // stream-K launch: the CTAs that shared this block's tile
            const long long u0 = (long long)(lb / R.tile_blks) * R.sk_S;
            ns = sk_owner(u0 + R.sk_S - 1, R.sk_U, R.sk_G) - sk_owner(u0, R.sk_U, R.sk_G) + 1;
        }
        for (int s = 0; s < ns; ++s) {
            const float *p = accp + (size_t)(R.slot0 + s) * n_iblk_shard * BLK_ELEMS + loff;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float4 v = *reinterpret_cast<const float4 *>(p + c * BLK);
                A[c].x += v.x; A[c].y += v.y; A[c].z += v.z; A[c].w += v.w;
            }
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        A[c].x *= acc_scale; A[c].y *= acc_scale; A[c].z *= acc_scale; A[c].w *= acc_scale;
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
    for (int d = 0; d < dests.n; ++d) {   // own next buffer, or every GPU's (P2P stores over NVLink)
        float *dst = reinterpret_cast<float *>(dests.p[d]) + goff;
#pragma unroll
        for (int c = 0; c < 4; ++c) *reinterpret_cast<float4 *>(dst + c * BLK) = P[c];
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) *reinterpret_cast<float4 *>(vel + loff + c * BLK) = V[c];
}

// With the cross-process exchange (`sig.n > 0`) the kernel is integrate + allgather + completion signal in one:
// the last CTA to finish publishes this rank's step counter into every peer's flag array.
__global__ void __launch_bounds__(256)
integrate_f32_kernel(const float *__restrict__ posm_cur, PeerDests dests, PeerSignal sig,
                     float *__restrict__ vel, float *__restrict__ acc,
                     const float *__restrict__ accp, float acc_scale, SlotPlan slots, int i_blk0,
                     int n_iblk_shard, int acc_only, long long n_real, IntegParams ip)
{
    integrate_f32_group(posm_cur, dests, vel, acc, accp, acc_scale, slots, i_blk0, n_iblk_shard, acc_only, n_real, ip);
    signal_peers_when_grid_done(sig);
}

cudaError_t launch_integrate_f32(const IntegLaunch &L, cudaStream_t st)
{
    const int threads = L.n_iblk_shard * 64;
    const int grid = (threads + 255) / 256;
    integrate_f32_kernel<<<grid, 256, 0, st>>>((const float *)L.posm_cur, L.dests, L.signal,
                                               (float *)L.vel, (float *)L.acc,
                                               (const float *)L.accp, L.acc_scale, L.slots, L.i_blk0,
                                               L.n_iblk_shard, L.acc_only, L.n_real, L.ip);
    return cudaGetLastError();
}

} // namespace nb
