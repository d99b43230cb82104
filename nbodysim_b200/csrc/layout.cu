// layout.cu -- body storage conversion (host AoS <-> device blocked SoA) and fp64 diagnostics.
//
// pack/unpack restate nothing arithmetic: they move the reference's 64-byte `Body` records
// (Body.hpp:6-14) into the device layout described in common.cuh and back
// (`SHARED_BODIES = simulation->bodies`, main.cpp:625, is the reference-side analogue).
#include "kernels.h"
#include "../../include/nbody_body.h"

namespace nb {

// `aos` holds the records [aos_first, ...) of the caller's array: the whole array (aos_first = 0, first = 0, count =
// n_padded: nbody_gpu_init) or only this rank's shard (aos_first = first = shard_start, count = shard_count: uploads
// of a distributed context).  `check_mass` != 0: every real body must still have exactly this mass (the uniform-mass
// force kernel is in use); a violation raises status[2].
template <typename T>
__global__ void __launch_bounds__(256)
pack_kernel(const nbody_body_t *__restrict__ aos, size_t aos_first, size_t first, size_t count, size_t n, size_t shard_start,
            size_t shard_count, T *__restrict__ posm, T *__restrict__ vel, T *__restrict__ acc, int dims,
            float check_mass, unsigned *status)
{
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const size_t i = first + k;
    // each thread reads its 64-byte record as four 16-byte vectors (a warp covers 2 KiB contiguous)
    // zero-mass padding beyond n sits far away (1e18): it contributes exactly 0 to every sum, also in
    // the uniform-mass kernel, where (r^2)^-3/2 underflows to 0 instead of being multiplied by m = 0
    float4 p = make_float4(PAD_POS, PAD_POS, PAD_POS, 0), v = make_float4(0, 0, 0, 0), a = v, mr = v;
    if (i < n) {
        const float4 *r = reinterpret_cast<const float4 *>(aos + (i - aos_first));
        p = r[0]; v = r[1]; a = r[2]; mr = r[3];
        if (check_mass != 0.f && mr.x != check_mass && status) *reinterpret_cast<volatile unsigned *>(status + 2) = 1u;
        // 2-D callers (the reference) leave the Vec2 tail padding indeterminate: never read z from it
        if (dims == 2) { p.z = 0.f; v.z = 0.f; a.z = 0.f; }
    }
    const size_t g = blk_index(i, 0);
    posm[g] = (T)p.x; posm[g + BLK] = (T)p.y; posm[g + 2 * BLK] = (T)p.z;
    posm[g + 3 * BLK] = (T)mr.x;
    if (i >= shard_start && i < shard_start + shard_count) {
        const size_t l = blk_index(i - shard_start, 0);
        vel[l] = (T)v.x; vel[l + BLK] = (T)v.y; vel[l + 2 * BLK] = (T)v.z; vel[l + 3 * BLK] = (T)mr.y;
        acc[l] = (T)a.x; acc[l + BLK] = (T)a.y; acc[l + 2 * BLK] = (T)a.z; acc[l + 3 * BLK] = (T)0;
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
unpack_kernel(nbody_body_t *__restrict__ aos, size_t n, size_t shard_start, size_t shard_count,
              const T *__restrict__ posm, const T *__restrict__ vel, const T *__restrict__ acc)
{
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; // index inside the shard
    const size_t i = shard_start + k;
    if (k >= shard_count || i >= n) return;
    const size_t g = blk_index(i, 0), l = blk_index(k, 0);
    float4 *r = reinterpret_cast<float4 *>(aos + k);              // staging holds the shard only
    r[0] = make_float4((float)posm[g], (float)posm[g + BLK], (float)posm[g + 2 * BLK], 0.f);
    r[1] = make_float4((float)vel[l], (float)vel[l + BLK], (float)vel[l + 2 * BLK], 0.f);
    r[2] = make_float4((float)acc[l], (float)acc[l + BLK], (float)acc[l + 2 * BLK], 0.f);
    r[3] = make_float4((float)posm[g + 3 * BLK], (float)vel[l + 3 * BLK], 0.f, 0.f);
}

template <typename T>
__global__ void __launch_bounds__(256)
unpack_f64_kernel(double *__restrict__ pos3, double *__restrict__ vel3, double *__restrict__ acc3,
                  size_t n, size_t shard_start, size_t shard_count, const T *__restrict__ posm,
                  const T *__restrict__ vel, const T *__restrict__ acc)
{
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t i = shard_start + k;
    if (k >= shard_count || i >= n) return;
    const size_t g = blk_index(i, 0), l = blk_index(k, 0);
    if (pos3) { pos3[3 * k] = posm[g]; pos3[3 * k + 1] = posm[g + BLK]; pos3[3 * k + 2] = posm[g + 2 * BLK]; }
    if (vel3) { vel3[3 * k] = vel[l]; vel3[3 * k + 1] = vel[l + BLK]; vel3[3 * k + 2] = vel[l + 2 * BLK]; }
    if (acc3) { acc3[3 * k] = acc[l]; acc3[3 * k + 1] = acc[l + BLK]; acc3[3 * k + 2] = acc[l + 2 * BLK]; }
}

cudaError_t launch_pack(const void *aos, size_t aos_first, size_t first, size_t count, size_t n, size_t shard_start,
                        size_t shard_count, void *posm, void *vel, void *acc, bool f64, int dims,
                        float check_mass, unsigned *status, cudaStream_t st)
{
    if (count == 0) return cudaSuccess;
    const unsigned grid = (unsigned)((count + 255) / 256);
    if (f64)
        pack_kernel<double><<<grid, 256, 0, st>>>((const nbody_body_t *)aos, aos_first, first, count, n, shard_start,
                                                  shard_count, (double *)posm, (double *)vel,
                                                  (double *)acc, dims, check_mass, status);
    else
        pack_kernel<float><<<grid, 256, 0, st>>>((const nbody_body_t *)aos, aos_first, first, count, n, shard_start,
                                                 shard_count, (float *)posm, (float *)vel,
                                                 (float *)acc, dims, check_mass, status);
    return cudaGetLastError();
}

cudaError_t launch_unpack(void *aos, size_t n, size_t shard_start, size_t shard_count,
                          const void *posm, const void *vel, const void *acc, bool f64,
                          cudaStream_t st)
{
    const unsigned grid = (unsigned)((shard_count + 255) / 256);
    if (f64)
        unpack_kernel<double><<<grid, 256, 0, st>>>((nbody_body_t *)aos, n, shard_start, shard_count,
                                                    (const double *)posm, (const double *)vel,
                                                    (const double *)acc);
    else
        unpack_kernel<float><<<grid, 256, 0, st>>>((nbody_body_t *)aos, n, shard_start, shard_count,
                                                   (const float *)posm, (const float *)vel,
                                                   (const float *)acc);
    return cudaGetLastError();
}

cudaError_t launch_unpack_f64(double *pos3, double *vel3, double *acc3, size_t n,
                              size_t shard_start, size_t shard_count, const void *posm,
                              const void *vel, const void *acc, bool f64, cudaStream_t st)
{
    const unsigned grid = (unsigned)((shard_count + 255) / 256);
    if (f64)
        unpack_f64_kernel<double><<<grid, 256, 0, st>>>(pos3, vel3, acc3, n, shard_start, shard_count,
                                                        (const double *)posm, (const double *)vel,
                                                        (const double *)acc);
    else
        unpack_f64_kernel<float><<<grid, 256, 0, st>>>(pos3, vel3, acc3, n, shard_start, shard_count,
                                                       (const float *)posm, (const float *)vel,
                                                       (const float *)acc);
    return cudaGetLastError();
}

// ---- energy / momentum diagnostic -----------------------------------------------------------------
// K = sum 1/2 m v^2, P = sum m v over the shard; W2 = - sum_{i in shard} sum_{j != i} m_i m_j
// rsqrt(r^2 + eps^2) over ALL sources (each pair counted twice over the whole system; the caller
// halves after summing ranks).  Sources are staged block by block through shared memory.  fp64 state:
// everything in double.  fp32 state: pair terms in fp32 (positions are floats; rsqrtf <= 2 ulp), summed
// in fp32 over one 256-source block and then accumulated in double, so the result resolves relative
// energy changes of ~1e-8 at 1/20 of the fp64 cost.
template <typename T>
__global__ void __launch_bounds__(256)
energy_kernel(const T *__restrict__ posm, const T *__restrict__ vel, size_t n_padded,
              size_t shard_start, size_t shard_count, double eps2, double *__restrict__ out5)
{
    __shared__ T tile[BLK_ELEMS];
    __shared__ double red[5][8];
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = k < shard_count;
    const size_t i = shard_start + (active ? k : 0);
    const size_t g = blk_index(i, 0), l = blk_index(active ? k : 0, 0);
    const T xi = posm[g], yi = posm[g + BLK], zi = posm[g + 2 * BLK];
    const double mi = active ? (double)posm[g + 3 * BLK] : 0.0;
    const double vx = vel[l], vy = vel[l + BLK], vz = vel[l + 2 * BLK];
    double K = 0.5 * mi * (vx * vx + vy * vy + vz * vz);
    double P0 = mi * vx, P1 = mi * vy, P2 = mi * vz;
    double w = 0;
    const T e2 = (T)eps2;
    const size_t nblk = n_padded / BLK;
    for (size_t b = 0; b < nblk; ++b) {
        __syncthreads();
        const T *src = posm + b * BLK_ELEMS;
#pragma unroll
        for (int c = 0; c < 4; ++c) tile[c * BLK + threadIdx.x] = src[c * BLK + threadIdx.x];
        __syncthreads();
        const long long self = (long long)i - (long long)(b * BLK);   // lane of the target inside this block, if any
        T wb = 0;
#pragma unroll 8
        for (int j = 0; j < BLK; ++j) {
            const T dx = tile[j] - xi, dy = tile[BLK + j] - yi, dz = tile[2 * BLK + j] - zi;
            const T r2 = dx * dx + dy * dy + dz * dz + e2;
            const T t = tile[3 * BLK + j] * (sizeof(T) == 8 ? (T)rsqrt((double)r2) : (T)rsqrtf((float)r2));
            wb += (j == self) ? (T)0 : t;
        }
        w += (double)wb;
    }
    double W = -mi * w;
    double v[5] = {K, W, P0, P1, P2};
#pragma unroll
    for (int c = 0; c < 5; ++c) {
        for (int o = 16; o > 0; o >>= 1) v[c] += __shfl_xor_sync(0xffffffffu, v[c], o);
        if ((threadIdx.x & 31) == 0) red[c][threadIdx.x >> 5] = v[c];
    }
    __syncthreads();
    if (threadIdx.x < 5) {
        double s = 0;
        for (int q = 0; q < 8; ++q) s += red[threadIdx.x][q];
        atomicAdd(&out5[threadIdx.x], s);
    }
}

cudaError_t launch_energy(const void *posm, const void *vel, size_t n_padded, size_t shard_start,
                          size_t shard_count, double eps2, bool f64, double *out5, cudaStream_t st)
{
    const unsigned grid = (unsigned)((shard_count + 255) / 256);
    if (f64)
        energy_kernel<double><<<grid, 256, 0, st>>>((const double *)posm, (const double *)vel,
                                                    n_padded, shard_start, shard_count, eps2, out5);
    else
        energy_kernel<float><<<grid, 256, 0, st>>>((const float *)posm, (const float *)vel, n_padded,
                                                   shard_start, shard_count, eps2, out5);
    return cudaGetLastError();
}

// ---- cross-process exchange: wait until every peer's pushes of the previous step have landed ----------------
// One lane per peer polls that peer's counter in this rank's own flag array (written by the peer's integrator
// kernel through its IPC mapping, st.release.sys) with ld.acquire.sys.  The launches that follow in the stream
// read the pushed positions.  A peer that stops advancing must not hang the GPU: after timeout_ns the kernel
// gives up and raises *status, which the host turns into NBODY_ESTATE at the next synchronisation.
__global__ void wait_peer_flags_kernel(const unsigned long long *flags, int world, int self, unsigned long long need,
                                       unsigned long long timeout_ns, unsigned *status)
{
    const int r = threadIdx.x;
    if (r >= world || r == self) return;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    unsigned spins = 0;
    while (ld_acquire_sys_u64(flags + r) < need) {
        __nanosleep(64);
        if ((++spins & 1023u) == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (t - t0 > timeout_ns) { atomicExch(status, 1u); return; }
        }
    }
}

cudaError_t launch_wait_peer_flags(const unsigned long long *flags, int world, int self, unsigned long long need,
                                   unsigned long long timeout_ns, unsigned *status, cudaStream_t st)
{
    wait_peer_flags_kernel<<<1, 32, 0, st>>>(flags, world, self, need, timeout_ns, status);
    return cudaGetLastError();
}

} // namespace nb
