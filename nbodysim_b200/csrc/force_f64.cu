// force_f64.cu -- tolerance-check mode: the same all-pairs sum and kick-drift with state and
// arithmetic in double (DFMA pipe, IEEE rsqrt).  Used to bound the fp32 kernels' error against an
// exact-math evaluation at sizes the CPU oracle cannot reach; never the headline number.
// Restates Quadtree.hpp:133-144 with rsqrt(x) in place of fast_inv_sqrt, and Body.hpp:34-38.
#include "kernels.h"

namespace nb {

constexpr int D_STAGE_BLKS = 1;                       // one block (8 KiB of doubles) per stage
constexpr int D_NSTAGE = 4;
constexpr int D_STAGE_ELEMS = D_STAGE_BLKS * BLK_ELEMS;
constexpr int D_STAGE_BYTES = D_STAGE_ELEMS * 8;
constexpr size_t F64_SMEM = (size_t)D_NSTAGE * D_STAGE_BYTES + 2 * D_NSTAGE * sizeof(uint64_t);

__device__ __forceinline__ void integrate_body_f64(double &px, double &py, double &pz, double &vx,
                                                   double &vy, double &vz, double ax, double ay,
                                                   double az, const IntegParams &ip)
{
    const double dt = (double)ip.dt;
    vx += ax * dt; vy += ay * dt; vz += az * dt;
    if (ip.flags & 1u) {
        const double v2 = vx * vx + vy * vy + vz * vz;
        const double mv = (double)ip.max_velocity;
        if (v2 > mv * mv) {
            const double scale = mv / sqrt(v2);
            vx *= scale; vy *= scale; vz *= scale;
        }
    }
    if (ip.flags & 2u) {
        const double d2 = px * px + py * py + pz * pz;
        const double sb = (double)ip.soft_boundary;
        if (d2 > sb * sb) {
            const double dist = sqrt(d2);
            const double force = (double)ip.boundary_force * exp(dist / sb - 1.0);
            const double k = -1.0 / dist, fd = force * dt;
            vx += px * k * fd; vy += py * k * fd; vz += pz * k * fd;
            vx *= (double)ip.damping; vy *= (double)ip.damping; vz *= (double)ip.damping;
        }
    }
    px += vx * dt; py += vy * dt; pz += vz * dt;
}

__global__ void __launch_bounds__(F64_THREADS, 4)
force_f64_kernel(const double *__restrict__ posm, double *__restrict__ accp, int i_blk0,
                 int i_blk_local0, int n_iblk_shard, int j_blk0, int j_nblk, int splits, int slot0,
                 double eps2)
{
    constexpr int I = F64_I;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *stage = reinterpret_cast<double *>(smem_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + (size_t)D_NSTAGE * D_STAGE_BYTES);
    uint64_t *empty = full + D_NSTAGE;
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < D_NSTAGE; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], F64_THREADS / 32);
        }
        mbar_fence_init();
    }
    __syncthreads();

    const int tile = blockIdx.x / splits;
    const int split = blockIdx.x - tile * splits;
    const int jb0 = j_blk0 + (int)(((long long)j_nblk * split) / splits);
    const int jb1 = j_blk0 + (int)(((long long)j_nblk * (split + 1)) / splits);
    const int nst = jb1 - jb0; // one block per stage
    const double *src = posm + (size_t)jb0 * BLK_ELEMS;

    auto issue = [&](int t) {
        const int s = t % D_NSTAGE;
        mbar_expect_tx(&full[s], (uint32_t)D_STAGE_BYTES);
        tma_bulk_g2s(stage + (size_t)s * D_STAGE_ELEMS, src + (size_t)t * BLK_ELEMS,
                     (uint32_t)D_STAGE_BYTES, &full[s]);
    };
    if (tid == 0)
        for (int t = 0; t < min(D_NSTAGE, nst); ++t) issue(t);

    double xi[I], yi[I], zi[I], ax[I], ay[I], az[I];
#pragma unroll
    for (int k = 0; k < I; ++k) {
        const double *b = posm + (size_t)(i_blk0 + tile) * BLK_ELEMS + k * F64_THREADS + tid;
        xi[k] = b[0]; yi[k] = b[BLK]; zi[k] = b[2 * BLK];
        ax[k] = ay[k] = az[k] = 0.0;
    }

    for (int t = 0; t < nst; ++t) {
        const int s = t % D_NSTAGE;
        mbar_wait(&full[s], (uint32_t)(t / D_NSTAGE) & 1u);
        const double *sx = stage + (size_t)s * D_STAGE_ELEMS;
#pragma unroll 2
        for (int j = 0; j < BLK; j += 2) {
            const double2 X = *reinterpret_cast<const double2 *>(sx + j);
            const double2 Y = *reinterpret_cast<const double2 *>(sx + BLK + j);
            const double2 Z = *reinterpret_cast<const double2 *>(sx + 2 * BLK + j);
            const double2 M = *reinterpret_cast<const double2 *>(sx + 3 * BLK + j);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const double xj = h ? X.y : X.x, yj = h ? Y.y : Y.x, zj = h ? Z.y : Z.x,
                             mj = h ? M.y : M.x;
#pragma unroll
                for (int k = 0; k < I; ++k) {
                    const double dx = xj - xi[k], dy = yj - yi[k], dz = zj - zi[k];
                    const double r2 = dx * dx + dy * dy + dz * dz;
                    const double ri = (r2 > 0.0) ? rsqrt(r2 + eps2) : 0.0; // Quadtree.hpp:139 guard
                    const double sgm = mj * ri * ri * ri;
                    ax[k] += dx * sgm; ay[k] += dy * sgm; az[k] += dz * sgm;
                }
            }
        }
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&empty[s]);
        if (tid == 0 && t >= 1 && (t - 1 + D_NSTAGE) < nst) {
            const int tp = t - 1;
            mbar_wait(&empty[tp % D_NSTAGE], (uint32_t)(tp / D_NSTAGE) & 1u);
            issue(tp + D_NSTAGE);
        }
    }
#pragma unroll
    for (int k = 0; k < I; ++k) {
        const int lb = i_blk_local0 + tile;
        double *o = accp + ((size_t)(slot0 + split) * n_iblk_shard + lb) * BLK_ELEMS +
                    k * F64_THREADS + tid;
        o[0] = ax[k]; o[BLK] = ay[k]; o[2 * BLK] = az[k];
    }
}

cudaError_t launch_force_f64(const ForceLaunch &L, cudaStream_t st)
{
    if (L.splits < 1 || L.j_nblk < L.splits) return cudaErrorInvalidValue;
    static_assert(F64_SMEM <= 48 * 1024, "above 48 KiB the per-device opt-in attribute would be needed");
    const int grid = (L.n_iblk / F64_TILE_BLKS) * L.splits;
    force_f64_kernel<<<grid, F64_THREADS, F64_SMEM, st>>>(
        (const double *)L.posm, (double *)L.accp, L.i_blk0, L.i_blk_local0, L.n_iblk_shard,
        L.j_blk0, L.j_nblk, L.splits, L.slot0, L.eps2_f64);
    return cudaGetLastError();
}

__device__ __forceinline__ void
integrate_f64_one(const double *__restrict__ posm_cur, const PeerDests &dests,
                  double *__restrict__ vel, double *__restrict__ acc,
                  const double *__restrict__ accp, float acc_scale, int nslots, int i_blk0,
                  int n_iblk_shard, int acc_only, long long n_real, const IntegParams &ip)
{
    const int gid = blockIdx.x * blockDim.x + threadIdx.x; // one body
    const int lb = gid >> 8;
    if (lb >= n_iblk_shard) return;
    const int q = gid & 255;
    const size_t loff = (size_t)lb * BLK_ELEMS + q;
    const size_t goff = (size_t)(i_blk0 + lb) * BLK_ELEMS + q;
    double a[3] = {0, 0, 0};
    for (int s = 0; s < nslots; ++s) {
        const double *p = accp + (size_t)s * n_iblk_shard * BLK_ELEMS + loff;
        a[0] += p[0]; a[1] += p[BLK]; a[2] += p[2 * BLK];
    }
    const double G = (double)acc_scale;
    a[0] *= G; a[1] *= G; a[2] *= G;
    acc[loff] = a[0]; acc[loff + BLK] = a[1]; acc[loff + 2 * BLK] = a[2];
    if (acc_only || (long long)(i_blk0 + lb) * BLK + q >= n_real) return;
    double px = posm_cur[goff], py = posm_cur[goff + BLK], pz = posm_cur[goff + 2 * BLK];
    const double m = posm_cur[goff + 3 * BLK];
    double vx = vel[loff], vy = vel[loff + BLK], vz = vel[loff + 2 * BLK];
    integrate_body_f64(px, py, pz, vx, vy, vz, a[0], a[1], a[2], ip);
    for (int d = 0; d < dests.n; ++d) {   // own next buffer, or every GPU's (P2P stores over NVLink)
        double *dst = reinterpret_cast<double *>(dests.p[d]) + goff;
        dst[0] = px; dst[BLK] = py; dst[2 * BLK] = pz; dst[3 * BLK] = m;
    }
    vel[loff] = vx; vel[loff + BLK] = vy; vel[loff + 2 * BLK] = vz;
}

__global__ void __launch_bounds__(256)
integrate_f64_kernel(const double *__restrict__ posm_cur, PeerDests dests, PeerSignal sig,
                     double *__restrict__ vel, double *__restrict__ acc,
                     const double *__restrict__ accp, float acc_scale, int nslots, int i_blk0,
                     int n_iblk_shard, int acc_only, long long n_real, IntegParams ip)
{
    integrate_f64_one(posm_cur, dests, vel, acc, accp, acc_scale, nslots, i_blk0, n_iblk_shard, acc_only, n_real, ip);
    signal_peers_when_grid_done(sig);   // cross-process exchange: see integrate_f32_kernel
}

cudaError_t launch_integrate_f64(const IntegLaunch &L, cudaStream_t st)
{
    int nslots = 0;                                       // fp64 launches are split launches: every slot of every range is filled
    for (int r = 0; r < L.slots.n; ++r) nslots += L.slots.r[r].nslots;
    const int threads = L.n_iblk_shard * BLK;
    integrate_f64_kernel<<<(threads + 255) / 256, 256, 0, st>>>(
        (const double *)L.posm_cur, L.dests, L.signal, (double *)L.vel, (double *)L.acc,
        (const double *)L.accp, L.acc_scale, nslots, L.i_blk0, L.n_iblk_shard, L.acc_only, L.n_real, L.ip);
    return cudaGetLastError();
}

} // namespace nb
