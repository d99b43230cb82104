// cluster_prims.cuh -- primitives for kernels that run as ONE thread-block cluster (8 CTAs, 16 where the device
// allows it) and keep a whole small scene inside it: the hardware cluster barrier (~0.2 us) replaces kernel
// boundaries (~3-9 us each at these sizes), and the CTAs exchange their partial results through distributed
// shared memory (mapa + ld.shared::cluster).
//
// At the reference's size (25,000 bodies) a Barnes-Hut build is ~20 dependent phases (8 radix-sort passes among
// them) over 0.3 MB of keys: every phase is latency-, not throughput-bound, so 16 SMs with cheap barriers beat
// 148 SMs with launch boundaries (profiles/r2_*: the multi-kernel build spends 71 of its 166 us in sort launches).
// Data stays in global memory (L2-resident); only counters, ranks and per-CTA totals live in shared memory.
#pragma once
#include "common.cuh"

namespace nb {

constexpr int CL_THREADS = 512, CL_WARPS = CL_THREADS / 32;
constexpr unsigned CL_MAX_CHUNK = 49152;          // sort items per CTA when their ranks are kept in shared memory (16-bit words: 96 KB)

__device__ __forceinline__ unsigned cl_rank()
{
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ unsigned cl_size()
{
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
// cluster-wide barrier; release / acquire at cluster scope: global and shared-memory writes made before it by any
// thread of the cluster are visible to every thread after it (the acquire also drops stale L1 lines)
__device__ __forceinline__ void cl_sync()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// read a 32-bit word of CTA `cta`'s shared memory, given the address of the same variable in this CTA
__device__ __forceinline__ unsigned cl_ld_u32(const unsigned *local, unsigned cta)
{
    unsigned remote, v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(local)), "r"(cta));
    asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(remote) : "memory");
    return v;
}

struct ClSmem {                                   // shared memory of a cluster kernel (~19 KB, + the ranks when they live on chip)
    unsigned counts[CL_WARPS][256];               // per-warp digit counts of a sort pass, then the warps' bases inside the digit
    unsigned hist[256];                           // this CTA's digit counts (read by the other CTAs)
    unsigned dest[256];                           // first output index for (this CTA, digit)
    unsigned xchg[64];                            // small per-CTA results exposed to the cluster (scan totals, box, maxima)
    unsigned warp_sums[CL_WARPS];
    unsigned misc[16];
};

// exclusive scan of one value per thread over the CL_THREADS threads of the CTA
__device__ __forceinline__ unsigned cl_block_excl_scan(unsigned v, unsigned *warp_sums, unsigned *total)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    __syncthreads();                              // warp_sums may still be read from a previous scan
    if (lane == 31) warp_sums[w] = incl;
    __syncthreads();
    unsigned before = 0, all = 0;
#pragma unroll
    for (int q = 0; q < CL_WARPS; ++q) {
        const unsigned s = warp_sums[q];
        if (q < w) before += s;
        all += s;
    }
    if (total) *total = all;
    return before + incl - v;
}

// sum over the cluster of one value per CTA (`mine`, significant on thread 0) and the sum over the lower-ranked
// CTAs; every thread receives both.  Contains one cluster barrier before the exchange and none after: the caller
// must pass another cl_sync() before `slot` is reused.
__device__ __forceinline__ void cl_exchange_sum(ClSmem &sm, unsigned *slot, unsigned mine, unsigned &before, unsigned &all)
{
    if (threadIdx.x == 0) *slot = mine;
    cl_sync();
    if (threadIdx.x < 32) {
        const unsigned lane = threadIdx.x, nc = cl_size(), rank = cl_rank();
        const unsigned v = lane < nc ? cl_ld_u32(slot, lane) : 0u;
        unsigned b = lane < rank ? v : 0u, a = v;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { b += __shfl_xor_sync(0xffffffffu, b, o); a += __shfl_xor_sync(0xffffffffu, a, o); }
        if (lane == 0) { sm.misc[0] = b; sm.misc[1] = a; }
    }
    __syncthreads();
    before = sm.misc[0];
    all = sm.misc[1];
    __syncthreads();
}

// Stable LSD radix sort (8-bit digits) of n (key [, value]) items in global memory by the whole cluster, over the key
// bits [begin_bit, end_bit).  Ping-pong between (ka, va) and (kb, vb); returns 0 when the result is in the a-buffers,
// 1 when it is in the b-buffers (a digit on which all keys agree costs no data movement and no flip).
// Per digit: every CTA ranks its contiguous chunk -- warps take contiguous sub-chunks row by row, __match_any_sync
// groups, one shared-memory atomic per group, so ranks follow input order (stable) -- publishes its 256 digit
// counts, reads the other CTAs' counts through distributed shared memory, and scatters.  Two cluster barriers per digit.
// `ranks` holds one RankT per item of this CTA's chunk (the item's rank among its warp's items of the same digit): shared
// memory (16-bit, chunks of up to CL_MAX_CHUNK items) or a global scratch array.  Every thread of the cluster must call it
// with the same arguments.
template <bool HAS_VALS, typename RankT>
__device__ int cl_radix_sort(ClSmem &sm, RankT *ranks, unsigned long long *ka, unsigned long long *kb, unsigned *va, unsigned *vb,
                             unsigned n, int begin_bit, int end_bit)
{
    const unsigned rank = cl_rank(), nc = cl_size();
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    const unsigned chunk = (n + nc - 1) / nc;
    const unsigned c0 = min(n, rank * chunk), c1 = min(n, c0 + chunk);
    const unsigned sub = ((((c1 - c0) + CL_WARPS - 1) / CL_WARPS) + 31u) & ~31u;    // items per warp: whole rows of 32
    const unsigned rows = sub / 32, w0 = c0 + (unsigned)w * sub;
    int flip = 0;
    for (int shift = begin_bit; shift < end_bit; shift += 8) {
#pragma unroll
        for (int q = 0; q < CL_WARPS * 256 / CL_THREADS; ++q) (&sm.counts[0][0])[q * CL_THREADS + tid] = 0;
        __syncthreads();
        for (unsigned r = 0; r < rows; ++r) {
            const unsigned i = w0 + r * 32 + lane;
            const bool valid = i < c1;
            const unsigned long long k = valid ? __ldcg(ka + i) : 0ull;
            const unsigned d = valid ? ((unsigned)(k >> shift) & 255u) : 256u + lane;     // invalid lanes match nobody
            const unsigned peers = __match_any_sync(0xffffffffu, d);
            const int leader = __ffs(peers) - 1;
            unsigned base = 0;
            if (valid && lane == leader) base = atomicAdd(&sm.counts[w][d], (unsigned)__popc(peers));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (valid) ranks[i - c0] = (RankT)(base + __popc(peers & ((1u << lane) - 1u)));
        }
        __syncthreads();
        if (tid < 256) {                           // digit tid: bases of the warps inside the digit, CTA total
            unsigned total = 0;
#pragma unroll
            for (int q = 0; q < CL_WARPS; ++q) { const unsigned c = sm.counts[q][tid]; sm.counts[q][tid] = total; total += c; }
            sm.hist[tid] = total;
        }
        cl_sync();
        unsigned tot = 0, before = 0;
        if (tid < 256)
            for (unsigned c = 0; c < nc; ++c) {
                const unsigned h = cl_ld_u32(&sm.hist[tid], c);
                tot += h;
                if (c < rank) before += h;
            }
        const int same = __syncthreads_or(tid < 256 && tot == n);       // every key has this digit: the pass is the identity
        const unsigned gstart = cl_block_excl_scan(tid < 256 ? tot : 0u, sm.warp_sums, nullptr);
        if (tid < 256) sm.dest[tid] = gstart + before;
        __syncthreads();
        if (!same) {
            for (unsigned r = 0; r < rows; ++r) {
                const unsigned i = w0 + r * 32 + lane;
                if (i < c1) {
                    const unsigned long long k = __ldcg(ka + i);
                    const unsigned d = (unsigned)(k >> shift) & 255u;
                    const unsigned dst = sm.dest[d] + sm.counts[w][d] + (unsigned)ranks[i - c0];
                    __stcg(kb + dst, k);
                    if (HAS_VALS) __stcg(vb + dst, __ldcg(va + i));
                }
            }
        }
        cl_sync();                                 // the scattered items are visible; hist / dest may be rewritten
        if (!same) {
            unsigned long long *tk = ka; ka = kb; kb = tk;
            unsigned *tv = va; va = vb; vb = tv;
            flip ^= 1;
        }
    }
    return flip;
}

// out[i] = in[0] + ... + in[i-1] for i in [0, n), by the whole cluster (contiguous chunk per CTA, contiguous run per
// thread); returns the grand total on every thread.  Ends with a cluster barrier (out[] is visible afterwards).
__device__ __forceinline__ unsigned cl_excl_scan(ClSmem &sm, const unsigned *in, unsigned *out, unsigned n)
{
    const unsigned rank = cl_rank(), nc = cl_size(), tid = threadIdx.x;
    const unsigned chunk = (n + nc - 1) / nc;
    const unsigned c0 = min(n, rank * chunk), c1 = min(n, c0 + chunk);
    const unsigned ipt = ((c1 - c0) + CL_THREADS - 1) / CL_THREADS;
    const unsigned t0 = min(c1, c0 + tid * ipt), t1 = min(c1, t0 + ipt);
    unsigned sum = 0;
    for (unsigned i = t0; i < t1; ++i) sum += __ldcg(in + i);
    unsigned total = 0;
    const unsigned excl = cl_block_excl_scan(sum, sm.warp_sums, &total);
    unsigned before, all;
    cl_exchange_sum(sm, &sm.xchg[0], total, before, all);
    unsigned run = before + excl;
    for (unsigned i = t0; i < t1; ++i) { const unsigned v = __ldcg(in + i); __stcg(out + i, run); run += v; }
    cl_sync();
    return all;
}

} // namespace nb
