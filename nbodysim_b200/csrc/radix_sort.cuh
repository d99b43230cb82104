// radix_sort.cuh -- hand-written stable LSD radix sort of 64-bit keys (+ optional 32-bit values) for the
// Barnes-Hut build (quadrant-path keys) and the collision pass (cell hashes, pair keys).
//
// 8 passes of 8 bits.  Per pass:
//   1. rs_hist_kernel    each CTA owns a contiguous tile of 4096 keys, each of its 8 warps a contiguous
//                        512-key chunk of the tile; digit counts per warp via __match_any_sync, summed per
//                        CTA into hist[digit][cta]                                  (digit-major)
//   2. rs_scan_kernel    one CTA: exclusive prefix sum over hist in digit-major order = the global start of
//                        every (digit, cta) bucket
//   3. rs_scatter_kernel re-reads the tile; rank of a key = start(digit, cta) + keys of the same digit in
//                        earlier warps of the CTA + earlier rows of this warp + earlier lanes of this row.
//                        Every term follows input order, so the sort is STABLE -- the Barnes-Hut merge of
//                        coincident bodies "in body-index order" (Quadtree::insert :56-60) depends on it.
// Ping-pong between two buffers; after the 8 passes the result is back in the FIRST buffer.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace nb {

constexpr int RS_THREADS = 256, RS_WARPS = RS_THREADS / 32, RS_ITEMS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;          // 4096 keys per CTA
constexpr int RS_CHUNK = RS_TILE / RS_WARPS;            // 512 keys per warp, contiguous
constexpr int RS_ROWS = RS_CHUNK / 32;                  // 16 rows of 32 keys per warp

__device__ __forceinline__ unsigned rs_digit(unsigned long long k, int shift) { return (unsigned)(k >> shift) & 255u; }

// counts[w][d]: how many keys of warp w's chunk have digit d
__device__ __forceinline__ void rs_warp_counts(const unsigned long long *__restrict__ keys, size_t n, size_t chunk0,
                                               int shift, unsigned (*counts)[256], int w, int lane)
{
    for (int d = lane; d < 256; d += 32) counts[w][d] = 0;
    __syncwarp();
    for (int r = 0; r < RS_ROWS; ++r) {
        const size_t i = chunk0 + (size_t)r * 32 + lane;
        const bool valid = i < n;
        const unsigned d = valid ? rs_digit(keys[i], shift) : 256u + lane;    // invalid lanes match nobody
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        if (valid && (peers & ((1u << lane) - 1u)) == 0u) counts[w][d] += __popc(peers);   // group leader
        __syncwarp();
    }
}

static __global__ void __launch_bounds__(RS_THREADS)
rs_hist_kernel(const unsigned long long *__restrict__ keys, size_t n, int shift, unsigned *__restrict__ hist, unsigned nblocks)
{
    __shared__ unsigned counts[RS_WARPS][256];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    rs_warp_counts(keys, n, (size_t)blockIdx.x * RS_TILE + (size_t)w * RS_CHUNK, shift, counts, w, lane);
    __syncthreads();
    unsigned s = 0;                                      // thread d sums digit d over the warps
#pragma unroll
    for (int q = 0; q < RS_WARPS; ++q) s += counts[q][threadIdx.x];
    hist[(size_t)threadIdx.x * nblocks + blockIdx.x] = s;
}

// exclusive scan of `m` counters in place, one CTA of 1024 threads
static __global__ void __launch_bounds__(1024) rs_scan_kernel(unsigned *__restrict__ a, size_t m)
{
    __shared__ unsigned part[1024];
    const size_t per = (m + 1023) / 1024, lo = (size_t)threadIdx.x * per, hi = lo + per < m ? lo + per : m;
    unsigned s = 0;
    for (size_t i = lo; i < hi; ++i) s += a[i];
    part[threadIdx.x] = s;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {                 // Hillis-Steele inclusive scan of the partials
        const unsigned v = (threadIdx.x >= (unsigned)o) ? part[threadIdx.x - o] : 0u;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    unsigned run = threadIdx.x ? part[threadIdx.x - 1] : 0u;
    for (size_t i = lo; i < hi; ++i) { const unsigned v = a[i]; a[i] = run; run += v; }
}

template <bool HAS_VALS>
static __global__ void __launch_bounds__(RS_THREADS)
rs_scatter_kernel(const unsigned long long *__restrict__ keys, const unsigned *__restrict__ vals, size_t n, int shift,
                  const unsigned *__restrict__ offs, unsigned nblocks, unsigned long long *__restrict__ keys_out,
                  unsigned *__restrict__ vals_out)
{
    __shared__ unsigned counts[RS_WARPS][256];           // per-warp digit counts, then running bases
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t chunk0 = (size_t)blockIdx.x * RS_TILE + (size_t)w * RS_CHUNK;
    rs_warp_counts(keys, n, chunk0, shift, counts, w, lane);
    __syncthreads();
    {   // thread d: base of (digit d, warp q) = global start of (d, this CTA) + counts of earlier warps
        const int d = threadIdx.x;
        unsigned run = offs[(size_t)d * nblocks + blockIdx.x];
#pragma unroll
        for (int q = 0; q < RS_WARPS; ++q) { const unsigned c = counts[q][d]; counts[q][d] = run; run += c; }
    }
    __syncthreads();
    for (int r = 0; r < RS_ROWS; ++r) {
        const size_t i = chunk0 + (size_t)r * 32 + lane;
        const bool valid = i < n;
        unsigned long long k = 0;
        unsigned d = 256u + lane;
        if (valid) { k = keys[i]; d = rs_digit(k, shift); }
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        const unsigned before = __popc(peers & ((1u << lane) - 1u));
        unsigned base = 0;
        if (valid) base = counts[w][d];
        __syncwarp();
        if (valid) {
            const unsigned dst = base + before;
            keys_out[dst] = k;
            if (HAS_VALS) vals_out[dst] = vals[i];
            if (before == 0u) counts[w][d] = base + __popc(peers);      // group leader advances the running base
        }
        __syncwarp();
    }
}

inline size_t radix_sort_temp_bytes(size_t n) { return (((n + RS_TILE - 1) / RS_TILE) * 256 + 256) * sizeof(unsigned); }

// Stable sort of n (key, value) pairs by key.  Result ends in keys_a / vals_a (8 ping-pong passes).
// vals_a == nullptr sorts keys only.  `temp` holds radix_sort_temp_bytes(n).  begin_bit/end_bit (multiples
// of 8) restrict the passes when the caller knows which key bits can differ.
inline cudaError_t radix_sort_u64(unsigned long long *keys_a, unsigned long long *keys_b, unsigned *vals_a, unsigned *vals_b,
                                  size_t n, void *temp, cudaStream_t st, int begin_bit = 0, int end_bit = 64, int *launches = nullptr)
{
    if (n == 0) return cudaSuccess;
    const unsigned nblocks = (unsigned)((n + RS_TILE - 1) / RS_TILE);
    unsigned *hist = (unsigned *)temp;
    unsigned long long *kin = keys_a, *kout = keys_b;
    unsigned *vin = vals_a, *vout = vals_b;
    int passes = 0;
    for (int shift = begin_bit; shift < end_bit; shift += 8) {
        rs_hist_kernel<<<nblocks, RS_THREADS, 0, st>>>(kin, n, shift, hist, nblocks);
        rs_scan_kernel<<<1, 1024, 0, st>>>(hist, (size_t)nblocks * 256);
        if (vals_a)
            rs_scatter_kernel<true><<<nblocks, RS_THREADS, 0, st>>>(kin, vin, n, shift, hist, nblocks, kout, vout);
        else
            rs_scatter_kernel<false><<<nblocks, RS_THREADS, 0, st>>>(kin, nullptr, n, shift, hist, nblocks, kout, nullptr);
        unsigned long long *tk = kin; kin = kout; kout = tk;
        unsigned *tv = vin; vin = vout; vout = tv;
        ++passes;
    }
    if (launches) *launches += 3 * passes;
    if (passes & 1) {   // odd number of passes: bring the result back to the first buffer
        cudaError_t e = cudaMemcpyAsync(keys_a, keys_b, n * 8, cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) return e;
        if (vals_a && (e = cudaMemcpyAsync(vals_a, vals_b, n * 4, cudaMemcpyDeviceToDevice, st)) != cudaSuccess) return e;
    }
    return cudaGetLastError();
}

} // namespace nb
