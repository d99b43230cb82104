// radix_sort.cuh -- hand-written stable LSD radix sort of 64-bit keys (+ optional 32-bit values) for the
// Barnes-Hut build (quadrant-path keys) and the collision pass (cell hashes, pair keys).
//
// Single-pass-per-digit design ("onesweep": one read and one write of the data per 8-bit digit):
//   os_hist_kernel   ONE read of the keys builds the digit histograms of all passes (shared-memory atomics,
//                    then one global atomic per (pass, digit) and CTA).
//   os_pass_kernel   per 8-bit digit.  A CTA takes the next tile of 4096 keys (ticket = start order, so a CTA
//                    only ever waits for tiles that are already running), ranks its keys by digit in
//                    registers (per-warp __match_any_sync, contiguous 512-key chunk per warp, rows in input
//                    order => STABLE), publishes its per-digit counts, and obtains the count of every digit
//                    in all EARLIER tiles by decoupled look-back over the tiles' status words (one thread per
//                    digit).  Keys (and values) are then permuted into digit order in shared memory and
//                    written out: consecutive threads write consecutive addresses within a digit's run.
//                    A digit on which all keys agree (unused high bytes) makes the pass a plain copy.
// HBM-/L2-bound: 12 B read + 12 B written per (key, value) pair and digit, plus 8 B per key for the histograms.
// Stability matters: the Barnes-Hut merge of coincident bodies "in body-index order" (Quadtree::insert
// :56-60) depends on it.  Ping-pong between two buffers; the result always ends in the FIRST buffer.
#pragma once
#include <cuda_runtime.h>
#include "common.cuh"                      // pdl_enter / launch_pdl
#include <cstdint>
#include <cstdlib>
#include <algorithm>

namespace nb {

constexpr int RS_THREADS = 256, RS_WARPS = RS_THREADS / 32;
// keys per thread: 16 (tile of 4096 keys, 512 contiguous keys = 16 rows of 32 per warp) for large inputs, 8 (tile
// of 2048) while the larger tile would leave SMs without a CTA
// keys per thread (rows of 32 keys per warp) by input size, measured on B200 (profiles/r2_sort_pass_trace.txt): small
// inputs are latency-bound, so more and smaller tiles win (25,000 pairs, 4 digits + repair: 29.6 us with 8 rows, 25.2 with 4,
// 23.8 with 2); large ones want long runs per digit in the scattered writes
constexpr int RS_ROWS_LARGE = 16, RS_ROWS_SMALL = 8;
constexpr size_t RS_ROWS2_MAX_N = 40000, RS_ROWS4_MAX_N = 200000;
#ifndef RS_SMALL_TILE_MAX_N
#define RS_SMALL_TILE_MAX_N (6u << 20)
#endif
constexpr int rs_rows_for(size_t n) { return n <= RS_ROWS2_MAX_N ? 2 : n <= RS_ROWS4_MAX_N ? 4 : n <= (size_t)RS_SMALL_TILE_MAX_N ? RS_ROWS_SMALL : RS_ROWS_LARGE; }
constexpr int RS_MAX_PASSES = 8;
// scratch layout: [RS_MAX_PASSES][256] digit histograms | RS_MISC_WORDS words (tickets of up to 2 x RS_MAX_PASSES pass
// launches, the grid-barrier word, the "tie runs too long" flag) | status words
constexpr int RS_MISC_WORDS = 64, RS_MISC_BARRIER = 2 * RS_MAX_PASSES, RS_MISC_FLAG = 2 * RS_MAX_PASSES + 1;
constexpr int RS_MISC_STAMPS = 20, RS_MISC_NSTAMPS = (RS_MISC_WORDS - RS_MISC_STAMPS) / 2;   // 64-bit time stamps of the all-passes kernel
// HIGH DIGITS FIRST (`lazy_low_bits` = L > 0).  Quadrant-path keys of distinct bodies almost always differ within their
// leading bits; the low digits only order the few bodies that share a deep cell.  So: LSD passes over the digits at and
// above bit L only, then ONE in-place repair: the first key of every run of keys that agree above bit L insertion-sorts
// its run by the full key (stable; runs are 2-3 keys in practice; one thread, ~1 us per element moved, so the limit is
// small).  A run longer than RS_TIE_RUN_MAX raises a flag and
// the full sort over all digits runs after all (same result: LSD passes are stable whatever order they start from).
constexpr int RS_TIE_RUN_MAX = 6;

__device__ __forceinline__ unsigned rs_digit(unsigned long long k, int shift) { return (unsigned)(k >> shift) & 255u; }

// exclusive scan of one value per thread over the 256 threads of the CTA (scratch: 8 words of shared memory)
__device__ __forceinline__ unsigned rs_block_excl_scan(unsigned v, unsigned *warp_sums, unsigned *total = nullptr)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    __syncthreads();                                      // warp_sums may still be read from a previous scan
    if (lane == 31) warp_sums[w] = incl;
    __syncthreads();
    unsigned before = 0, all = 0;
#pragma unroll
    for (int q = 0; q < RS_WARPS; ++q) {
        const unsigned s = warp_sums[q];
        if (q < w) before += s;
        all += s;
    }
    if (total) *total = all;
    return before + incl - v;
}

// For a kernel that PRODUCES the keys (256 threads per CTA): count the 8 digits of every key it writes into a
// shared histogram and add the CTA's counts to the sort's global histograms (the first 8 x 256 words of `temp`).
__device__ __forceinline__ void rs_hist_clear(unsigned (*sh)[256])
{
    for (int p = 0; p < RS_MAX_PASSES; ++p) sh[p][threadIdx.x] = 0;
    __syncthreads();
}
__device__ __forceinline__ void rs_hist_add_key(unsigned (*sh)[256], unsigned long long k)
{
#pragma unroll
    for (int p = 0; p < RS_MAX_PASSES; ++p) atomicAdd(&sh[p][(unsigned)(k >> (8 * p)) & 255u], 1u);
}
__device__ __forceinline__ void rs_hist_flush(unsigned (*sh)[256], unsigned *__restrict__ hist)
{
    __syncthreads();
    for (int p = 0; p < RS_MAX_PASSES; ++p) {
        const unsigned c = sh[p][threadIdx.x];
        if (c) atomicAdd(&hist[p * 256 + threadIdx.x], c);
    }
}

// hist[p][d] += number of keys whose digit of pass (pass0 + p) is d
static __global__ void __launch_bounds__(RS_THREADS)
os_hist_kernel(const unsigned long long *__restrict__ keys, size_t n, const unsigned *__restrict__ n_dev, int pass0, int npasses,
               unsigned *__restrict__ hist)
{
    if (n_dev) n = (*n_dev < n) ? *n_dev : n;             // live item count kept on the device (n is its bound)
    __shared__ unsigned sh[RS_MAX_PASSES][256];
    for (int p = 0; p < npasses; ++p) sh[p][threadIdx.x] = 0;
    __syncthreads();
    for (size_t i = (size_t)blockIdx.x * RS_THREADS + threadIdx.x; i < n; i += (size_t)gridDim.x * RS_THREADS) {
        const unsigned long long k = keys[i] >> (8 * pass0);
#pragma unroll
        for (int p = 0; p < RS_MAX_PASSES; ++p)
            if (p < npasses) atomicAdd(&sh[p][(unsigned)(k >> (8 * p)) & 255u], 1u);
    }
    __syncthreads();
    for (int p = 0; p < npasses; ++p) {
        const unsigned c = sh[p][threadIdx.x];
        if (c) atomicAdd(&hist[p * 256 + threadIdx.x], c);
    }
}

// status word of (tile, digit): [63:32] tag = 2 * (pass + 1) + is_inclusive_prefix, [31:0] count.  0 = not ready;
// words left over from an earlier pass carry another tag and read as "not ready" too.
__device__ __forceinline__ unsigned long long rs_pack(unsigned tag, unsigned count) { return ((unsigned long long)tag << 32) | count; }
// status words are single 64-bit words carrying flag and value together: relaxed device-scope accesses (served by L2)
__device__ __forceinline__ unsigned long long rs_ld_status(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void rs_st_status(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
constexpr int RS_LOOKBACK_WINDOW = 4;                     // predecessors' status words fetched at once

// One tile of one digit pass (the body of both kernels below).  `tile` < 0: take the next tile by ticket.  IDENTITY_COPY:
// a digit on which all keys agree makes the pass a plain copy (the launch-per-pass form keeps its ping-pong fixed);
// the all-passes kernel tests that itself and skips such a pass altogether.
// keys of digit `tid` in all tiles before `tile` (> 0): walk back over the tiles' status words, WINDOW at a time
template <int WINDOW>
__device__ __forceinline__ unsigned rs_lookback(const unsigned long long *status, unsigned tile, int tid, unsigned pass)
{
    unsigned excl = 0;
    long long t = (long long)tile - 1;
    bool done = false;
    while (!done) {
        unsigned long long v[WINDOW];
#pragma unroll
        for (int j = 0; j < WINDOW; ++j) {
            const long long tj = t - j > 0 ? t - j : 0;
            v[j] = rs_ld_status(status + (size_t)tj * 256 + tid);
        }
#pragma unroll
        for (int j = 0; j < WINDOW; ++j) {
            if (done) break;
            const long long tj = t - j > 0 ? t - j : 0;
            unsigned vt = (unsigned)(v[j] >> 32);
            while ((vt >> 1) != pass + 1u) {                               // not published yet: poll again
                v[j] = rs_ld_status(status + (size_t)tj * 256 + tid);
                vt = (unsigned)(v[j] >> 32);
            }
            excl += (unsigned)v[j];
            done = (vt & 1u) != 0u;                                        // an inclusive prefix ends the walk
        }
        t -= WINDOW;
    }
    return excl;
}

#ifdef RS_FINE_TRACE   // tools only: the LAST CTA stamps the global timer inside a pass (fine_stamps[] in global memory)
__device__ unsigned long long rs_fine_stamps[64];
__device__ int rs_fine_k;
#define RS_FINE() do { if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0 && rs_fine_k < 64) { unsigned long long t_; \
                       asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); rs_fine_stamps[rs_fine_k++] = t_; } } while (0)
#else
#define RS_FINE() do { } while (0)
#endif
template <bool HAS_VALS, int RS_ROWS, bool IDENTITY_COPY>
__device__ __forceinline__ void
os_pass_tile(const unsigned long long *__restrict__ keys, const unsigned *__restrict__ vals, size_t n, int shift, unsigned pass,
             const unsigned *__restrict__ hist_p, unsigned *__restrict__ ticket, int fixed_tile, unsigned long long *status,
             unsigned long long *__restrict__ keys_out, unsigned *__restrict__ vals_out, bool lb_wide = false)
{
    constexpr int RS_ITEMS = RS_ROWS, RS_CHUNK = 32 * RS_ROWS, RS_TILE = RS_THREADS * RS_ROWS;
    __shared__ unsigned long long skeys[RS_TILE];        // the tile in digit order; reused for the values afterwards
    unsigned *svals = reinterpret_cast<unsigned *>(skeys);
    __shared__ unsigned counts[RS_WARPS][256];           // per-warp digit counts, then the warps' bases inside the digit
    __shared__ unsigned dig_excl[256];                   // start of digit d in the tile's sorted order
    __shared__ unsigned gbase[256];                      // output index of sorted-local index i of digit d = gbase[d] + i
    __shared__ unsigned warp_sums[RS_WARPS];
    __shared__ unsigned s_tile;

    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    RS_FINE();   // 0: start
    __syncthreads();                                      // a previous tile / pass of this CTA may still read the shared arrays
    if (tid == 0) s_tile = fixed_tile >= 0 ? (unsigned)fixed_tile : atomicAdd(ticket, 1u);
#pragma unroll
    for (int q = 0; q < RS_WARPS; ++q) counts[q][tid] = 0;
    __syncthreads();
    const unsigned tile = s_tile;
    const size_t tile0 = (size_t)tile * RS_TILE;
    if (tile0 >= n) return;                               // grid sized for the bound: nothing left for this CTA
    const unsigned cnt = (unsigned)((n - tile0 < (size_t)RS_TILE) ? (n - tile0) : (size_t)RS_TILE);

    // ---- every key has the same digit (e.g. the unused high bytes of small keys): the pass is the identity
    if (IDENTITY_COPY && __syncthreads_or(hist_p[tid] == (unsigned)n)) {
#pragma unroll
        for (int j = 0; j < RS_ITEMS; ++j) {
            const unsigned i = (unsigned)j * RS_THREADS + tid;
            if (i < cnt) {
                keys_out[tile0 + i] = keys[tile0 + i];
                if (HAS_VALS) vals_out[tile0 + i] = vals[tile0 + i];
            }
        }
        return;
    }

    // ---- load the warp's 512-key chunk row by row and rank every key among the chunk's keys of the same digit
    unsigned long long k[RS_ROWS];
    unsigned rank[RS_ROWS];
    constexpr bool EARLY_VALS = HAS_VALS && RS_ROWS <= 8;   // registers permitting, fetch the values with the keys
    unsigned v[EARLY_VALS ? RS_ROWS : 1];
#pragma unroll
    for (int r = 0; r < RS_ROWS; ++r) {
        const unsigned i = (unsigned)w * RS_CHUNK + (unsigned)r * 32 + lane;
        k[r] = (i < cnt) ? keys[tile0 + i] : ~0ull;
        if (EARLY_VALS) v[r] = (i < cnt) ? vals[tile0 + i] : 0u;
    }
    // A group of lanes holding the same digit is ranked by its first lane, which advances the warp's counter of that
    // digit; rows in program order => stable.
#ifdef RS_FINE_TRACE
    if (__syncthreads_or(k[RS_ROWS - 1] == 1ull && v[0] == 7u)) return;       // tools only: all loads have landed
    RS_FINE();   // 0a: keys (and values) loaded
#endif
#pragma unroll
    for (int r = 0; r < RS_ROWS; ++r) {
        const unsigned i = (unsigned)w * RS_CHUNK + (unsigned)r * 32 + lane;
        const bool valid = i < cnt;
        const unsigned d = valid ? rs_digit(k[r], shift) : 0u;
        // lanes holding the same digit: eight ballots, one per digit bit (measured on B200: __match_any_sync on 32 mostly
        // distinct digits costs ~600 cycles per row, this ~60; profiles/r2_sort_pass_trace.txt)
        unsigned peers = __ballot_sync(0xffffffffu, valid);
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const unsigned bal = __ballot_sync(0xffffffffu, (d >> b) & 1u);
            peers &= ((d >> b) & 1u) ? bal : ~bal;
        }
        if (!valid) peers = 1u << lane;                                       // invalid lanes match nobody
        const int leader = __ffs(peers) - 1;
        unsigned base = 0;
        // the warp owns its row of counters and a digit has ONE leader per row: a plain read-modify-write, ordered across
        // the rows (whose leaders for a digit may be different lanes) by __syncwarp -- no shared-memory atomic
        if (valid && lane == leader) { base = counts[w][d]; counts[w][d] = base + (unsigned)__popc(peers); }
        __syncwarp();
        base = __shfl_sync(0xffffffffu, base, leader);
        rank[r] = base + __popc(peers & ((1u << lane) - 1u));
    }
    __syncthreads();
    RS_FINE();   // 1: keys loaded and ranked

    // ---- thread d: digit d's count in this tile, bases of the warps inside the digit
    unsigned total = 0;
#pragma unroll
    for (int q = 0; q < RS_WARPS; ++q) { const unsigned c = counts[q][tid]; counts[q][tid] = total; total += c; }
    const size_t sidx = (size_t)tile * 256 + tid;
    const unsigned tag = 2u * (pass + 1u);
    rs_st_status(status + sidx, rs_pack(tile == 0 ? tag + 1u : tag, total));                 // tile 0's aggregate IS its inclusive prefix

    const unsigned dexcl = rs_block_excl_scan(total, warp_sums);               // start of digit d in the sorted tile
    const unsigned gstart = rs_block_excl_scan(hist_p[tid], warp_sums);        // start of digit d in the whole output
    dig_excl[tid] = dexcl;

    RS_FINE();   // 2: published, scans done
    // ---- decoupled look-back: keys of digit d in all earlier tiles
    // The walk goes tile-1, tile-2, ... adding aggregates until it meets an inclusive prefix (tile 0 always
    // publishes one).  Each step is an L2 round trip, so RS_LOOKBACK_WINDOW predecessors are fetched at once
    // and consumed in order; a word that is not published yet is polled again.
    unsigned excl = 0;
    if (tile > 0) {
        excl = lb_wide ? rs_lookback<16>(status, tile, tid, pass) : rs_lookback<RS_LOOKBACK_WINDOW>(status, tile, tid, pass);
        rs_st_status(status + sidx, rs_pack(tag + 1u, excl + total));
    }
    gbase[tid] = gstart + excl - dexcl;
    __syncthreads();
    RS_FINE();   // 3: look-back done

    // ---- permute the keys into digit order in shared memory
    unsigned pos[RS_ROWS];
#pragma unroll
    for (int r = 0; r < RS_ROWS; ++r) {
        const unsigned i = (unsigned)w * RS_CHUNK + (unsigned)r * 32 + lane;
        pos[r] = 0;
        if (i < cnt) {
            const unsigned d = rs_digit(k[r], shift);
            pos[r] = dig_excl[d] + counts[w][d] + rank[r];
            skeys[pos[r]] = k[r];
        }
    }
    __syncthreads();

    // ---- write out: thread i, i + 256, ... of the sorted tile; runs of one digit are contiguous in the output
    unsigned dst[RS_ITEMS];
#pragma unroll
    for (int j = 0; j < RS_ITEMS; ++j) {
        const unsigned i = (unsigned)j * RS_THREADS + tid;
        dst[j] = 0;
        if (i < cnt) {
            const unsigned long long key = skeys[i];
            dst[j] = gbase[rs_digit(key, shift)] + i;
            keys_out[dst[j]] = key;
        }
    }
    RS_FINE();   // 4: keys permuted and stored
    if (HAS_VALS) {      // the same permutation for the values, through the same shared-memory buffer
        __syncthreads();
#pragma unroll
        for (int r = 0; r < RS_ROWS; ++r) {
            const unsigned i = (unsigned)w * RS_CHUNK + (unsigned)r * 32 + lane;
            if (i < cnt) svals[pos[r]] = EARLY_VALS ? v[r] : vals[tile0 + i];
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < RS_ITEMS; ++j) {
            const unsigned i = (unsigned)j * RS_THREADS + tid;
            if (i < cnt) vals_out[dst[j]] = svals[i];
        }
    }
}

// `run_if` (the tie-run flag of a high-digits-first sort): the pass belongs to the full sort that only runs when the
// repair gave up; a launch that finds the flag clear returns at once.
template <bool HAS_VALS, int RS_ROWS>
static __global__ void __launch_bounds__(RS_THREADS, (RS_ROWS <= 8 ? 4 : 3))
os_pass_kernel(const unsigned long long *__restrict__ keys, const unsigned *__restrict__ vals, size_t n, const unsigned *__restrict__ n_dev,
               int shift, unsigned pass,
               const unsigned *__restrict__ hist_p, unsigned *__restrict__ ticket, unsigned long long *status,
               unsigned long long *__restrict__ keys_out, unsigned *__restrict__ vals_out, const unsigned *run_if)
{
    if (run_if && *run_if == 0u) return;
    if (n_dev) n = (*n_dev < n) ? *n_dev : n;             // live item count kept on the device (n is its bound)
    os_pass_tile<HAS_VALS, RS_ROWS, true>(keys, vals, n, shift, pass, hist_p, ticket, -1, status, keys_out, vals_out);
}

// ---- repair after a high-digits-first sort ---------------------------------------------------------------------------------
// Element i: if it is the FIRST of a run of keys that agree above bit `low_bits`, insertion-sort the run in place by the
// full key (strict comparison => stable).  Runs are disjoint, so the leaders never touch each other's elements; everything
// goes through L2 (the keys were scattered by other SMs moments ago).  A run longer than RS_TIE_RUN_MAX is left alone and
// reported through `flag`.
template <bool HAS_VALS>
__device__ __forceinline__ void rs_repair_run(unsigned long long *keys, unsigned *vals, size_t n, int low_bits, size_t i, unsigned *flag)
{
    if (i + 1 >= n) return;
    const unsigned long long hi = __ldcg(keys + i) >> low_bits;
    if ((__ldcg(keys + i + 1) >> low_bits) != hi) return;
    if (i > 0 && (__ldcg(keys + i - 1) >> low_bits) == hi) return;           // not the first of its run
    size_t e = i + 2;
    while (e < n && e - i <= (size_t)RS_TIE_RUN_MAX && (__ldcg(keys + e) >> low_bits) == hi) ++e;
    if (e - i > (size_t)RS_TIE_RUN_MAX) { *reinterpret_cast<volatile unsigned *>(flag) = 1u; return; }
    for (size_t a = i + 1; a < e; ++a) {
        const unsigned long long ka = __ldcg(keys + a);
        const unsigned va = HAS_VALS ? __ldcg(vals + a) : 0u;
        size_t b = a;
        while (b > i) {
            const unsigned long long kb = __ldcg(keys + b - 1);
            if (!(kb > ka)) break;
            __stcg(keys + b, kb);
            if (HAS_VALS) __stcg(vals + b, __ldcg(vals + b - 1));
            --b;
        }
        if (b != a) {
            __stcg(keys + b, ka);
            if (HAS_VALS) __stcg(vals + b, va);
        }
    }
}

template <bool HAS_VALS>
static __global__ void __launch_bounds__(RS_THREADS)
os_repair_kernel(unsigned long long *keys, unsigned *vals, size_t n, const unsigned *__restrict__ n_dev, int low_bits, unsigned *flag)
{
    if (n_dev) n = (*n_dev < n) ? *n_dev : n;
    for (size_t i = (size_t)blockIdx.x * RS_THREADS + threadIdx.x; i < n; i += (size_t)gridDim.x * RS_THREADS)
        rs_repair_run<HAS_VALS>(keys, vals, n, low_bits, i, flag);
}

// ---- all digit passes in ONE kernel (small inputs: every tile co-resident) -------------------------------------------------
// At the reference's size a digit pass is 13 tiles of work and ~9 us of launch, ramp and drain: launched cooperatively
// with one CTA per tile, the passes follow one another inside the kernel, separated by a grid-wide barrier (the scattered
// keys of pass p must be complete before pass p + 1 reads them).  Same tile code, same decoupled look-back (the status
// words carry the pass in their tag).  A digit on which all keys agree is skipped without moving anything; the result
// is brought back to the first buffer pair at the end if an odd number of passes moved data.
__device__ __forceinline__ void rs_grid_barrier(unsigned *counter, unsigned target)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();                                   // this CTA's scattered keys are visible before it arrives
        atomicAdd(counter, 1u);
        while (*reinterpret_cast<volatile unsigned *>(counter) < target) { }
        __threadfence();
    }
    __syncthreads();
}

#ifndef RS_BOUNDS8
#define RS_BOUNDS8 2
#endif
template <bool HAS_VALS, int RS_ROWS>
static __global__ void __launch_bounds__(RS_THREADS, (RS_ROWS <= 4 ? 2 : RS_ROWS <= 8 ? RS_BOUNDS8 : 3))
os_sort_all_kernel(unsigned long long *keys_a, unsigned long long *keys_b, unsigned *vals_a, unsigned *vals_b, size_t n,
                   const unsigned *__restrict__ n_dev, int pass0, int npasses, int lazy_passes, const unsigned *__restrict__ hist,
                   unsigned *__restrict__ misc, unsigned long long *status)
{
    if (n_dev) n = (*n_dev < n) ? *n_dev : n;
    unsigned *barrier_word = misc + RS_MISC_BARRIER, *flag = misc + RS_MISC_FLAG;
    // tuning aid: CTA 0 leaves the global timer (ns) at the start, after every pass / the repair, and at the end in the
    // spare scratch words (read by tools/sort_check.cu --trace); a handful of instructions of one thread
    unsigned long long *stamps = reinterpret_cast<unsigned long long *>(misc + RS_MISC_STAMPS);
    int stamp_k = 0;
#define RS_STAMP() do { if (blockIdx.x == 0 && threadIdx.x == 0 && stamp_k < RS_MISC_NSTAMPS) { unsigned long long t_;                \
                        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); stamps[stamp_k++] = t_; } } while (0)
    RS_STAMP();
    unsigned long long *kin = keys_a, *kout = keys_b;
    unsigned *vin = vals_a, *vout = vals_b;
    unsigned syncs = 0, pass_id = 0;
    bool flipped = false;
    const size_t ntiles = (n + (size_t)RS_THREADS * RS_ROWS - 1) / ((size_t)RS_THREADS * RS_ROWS);
    // digits on which all keys agree (identity passes), for all passes at once: one round trip instead of one per pass
    __shared__ unsigned s_skip;
    if (threadIdx.x == 0) s_skip = 0;
    __syncthreads();
    {
        unsigned m = 0;
        for (int p = 0; p < npasses; ++p) m |= (hist[p * 256 + threadIdx.x] == (unsigned)n) ? (1u << p) : 0u;
        if (m) atomicOr(&s_skip, m);
    }
    __syncthreads();
    const unsigned skip = s_skip;
    // one CTA per tile: all tiles start a pass together, nobody finds an inclusive prefix early and the look-back runs
    // all the way to tile 0 -- fetch 16 predecessors per round trip instead of 4
    const bool wide = ntiles <= gridDim.x && ntiles <= 64;
    // lazy_passes > 0: only the digits from `lazy_passes` upwards first, then the repair, then -- if a tie run was too
    // long -- every digit after all
    for (int round = 0, first = lazy_passes; round < 2; ++round, first = 0) {
        for (int p = first; p < npasses; ++p, ++pass_id) {
            const unsigned *hist_p = hist + p * 256;
            if ((skip >> p) & 1u) continue;                                        // uniform over the grid: identity pass
            for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x)               // tiles in increasing order: look-back never waits on a later one
                os_pass_tile<HAS_VALS, RS_ROWS, false>(kin, vin, n, 8 * (pass0 + p), pass_id, hist_p, nullptr, (int)t, status, kout, vout, wide);
            RS_FINE();   // 5: values stored
            rs_grid_barrier(barrier_word, ++syncs * gridDim.x);
            RS_FINE();   // 6: grid barrier passed
            RS_STAMP();
            unsigned long long *tk = kin; kin = kout; kout = tk;
            unsigned *tv = vin; vin = vout; vout = tv;
            flipped = !flipped;
        }
        if (round == 1 || lazy_passes == 0) break;
        for (size_t i = (size_t)blockIdx.x * RS_THREADS + threadIdx.x; i < n; i += (size_t)gridDim.x * RS_THREADS)
            rs_repair_run<HAS_VALS>(kin, vin, n, 8 * (pass0 + lazy_passes), i, flag);
        rs_grid_barrier(barrier_word, ++syncs * gridDim.x);
        RS_STAMP();
        if (*reinterpret_cast<volatile unsigned *>(flag) == 0u) break;             // the usual case: sorted
    }
    if (flipped) {                                         // bring the result back to the first buffer pair
        for (size_t i = (size_t)blockIdx.x * RS_THREADS + threadIdx.x; i < n; i += (size_t)gridDim.x * RS_THREADS) {
            keys_a[i] = __ldcg(kin + i);
            if (HAS_VALS) vals_a[i] = __ldcg(vin + i);
        }
    }
    RS_STAMP();
#undef RS_STAMP
}

inline size_t radix_sort_temp_bytes(size_t n)
{
    // status words for the most tiles any input of up to n items is cut into (the tile grows with the input)
    size_t ntiles = 0;
    const size_t tier_max[4] = {RS_ROWS2_MAX_N, RS_ROWS4_MAX_N, (size_t)RS_SMALL_TILE_MAX_N, (size_t)-1};
    const int tier_rows[4] = {2, 4, RS_ROWS_SMALL, RS_ROWS_LARGE};
    for (int t = 0; t < 4; ++t) {
        const size_t m = n < tier_max[t] ? n : tier_max[t], tile = (size_t)RS_THREADS * tier_rows[t];
        ntiles = std::max(ntiles, (m + tile - 1) / tile);
    }
    return (size_t)(RS_MAX_PASSES * 256 + RS_MISC_WORDS) * sizeof(unsigned) + ntiles * 256 * sizeof(unsigned long long);
}

// CTAs the all-passes kernel may be launched with (all co-resident; a CTA takes tiles blockIdx, blockIdx + grid, ...);
// 0 disables it (NBODY_SORT_COOP=0, or a device without cooperative launch)
template <bool V> static inline const void *rs_all_kernel_for(int rows)
{
    return rows == 2 ? (const void *)os_sort_all_kernel<V, 2> : rows == 4 ? (const void *)os_sort_all_kernel<V, 4>
         : rows == RS_ROWS_SMALL ? (const void *)os_sort_all_kernel<V, RS_ROWS_SMALL> : (const void *)os_sort_all_kernel<V, RS_ROWS_LARGE>;
}
static inline int rs_coop_cta_limit(bool has_vals, int rows)
{
    static int limit[2][4] = {{-1, -1, -1, -1}, {-1, -1, -1, -1}};
    int &l = limit[has_vals ? 1 : 0][rows == 2 ? 0 : rows == 4 ? 1 : rows == RS_ROWS_SMALL ? 2 : 3];
    if (l < 0) {
        l = 0;
        const char *env = getenv("NBODY_SORT_COOP");
        int dev = 0, coop = 0, sms = 0, per_sm = 0;
        if (!(env && atoi(env) == 0) && cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev) == cudaSuccess && coop &&
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess) {
            const void *fn = has_vals ? rs_all_kernel_for<true>(rows) : rs_all_kernel_for<false>(rows);
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, RS_THREADS, 0) == cudaSuccess) l = sms * per_sm;
            else cudaGetLastError();
        }
    }
    return l;
}

// Stable sort of n (key, value) pairs by key.  Result ends in keys_a / vals_a.  vals_a == nullptr sorts keys only.
// `temp` holds radix_sort_temp_bytes(n).  begin_bit/end_bit (multiples of 8) restrict the passes when the caller
// knows which key bits can differ.  With `n_dev` the number of live items is read from device memory by the kernels
// (n is then only the bound the grids are sized for), so a producer's counter never has to travel to the host.
// `lazy_low_bits` (multiple of 8, 0 = off): high digits first, the bits below it only through the run repair (see
// RS_TIE_RUN_MAX above); honoured when the pass counts keep the ping-pong parity, silently a plain sort otherwise.
// Fully asynchronous on `st` (graph-capturable).
static inline cudaError_t radix_sort_u64(unsigned long long *keys_a, unsigned long long *keys_b, unsigned *vals_a, unsigned *vals_b,
                                  size_t n, void *temp, cudaStream_t st, int begin_bit = 0, int end_bit = 64, int *launches = nullptr,
                                  const unsigned *n_dev = nullptr, bool hist_ready = false, int lazy_low_bits = 0)
{
    if (n == 0 || end_bit <= begin_bit) return cudaSuccess;
    const int pass0 = begin_bit / 8, npasses = (end_bit - begin_bit + 7) / 8;
    if (begin_bit % 8 || npasses > RS_MAX_PASSES) return cudaErrorInvalidValue;
    static const bool lazy_off = getenv("NBODY_SORT_LAZY") && atoi(getenv("NBODY_SORT_LAZY")) == 0;
    int lazy_passes = (lazy_low_bits > begin_bit && lazy_low_bits < end_bit && lazy_low_bits % 8 == 0 && !lazy_off) ? (lazy_low_bits - begin_bit) / 8 : 0;
    const int rows = rs_rows_for(n);
    const size_t tile = (size_t)RS_THREADS * rows, ntiles = (n + tile - 1) / tile;
    unsigned *hist = (unsigned *)temp;                                    // [npasses][256]
    unsigned *misc = hist + RS_MAX_PASSES * 256;                          // tickets [2 x npasses], barrier word, tie-run flag
    unsigned long long *status = (unsigned long long *)(misc + RS_MISC_WORDS);
    // hist_ready: the caller zeroed `temp` and the producer of the keys already accumulated the digit histograms
    // of these passes into it (rs_hist_add_key / rs_hist_flush): one memset, one kernel and one read of the keys less
    cudaError_t e = cudaSuccess;
    if (!hist_ready && (e = cudaMemsetAsync(temp, 0, radix_sort_temp_bytes(n), st)) != cudaSuccess) return e;
    size_t hgrid = (n + 8 * RS_THREADS - 1) / (8 * RS_THREADS);             // >= 8 keys per thread, at most 8 CTAs per SM
    if (hgrid > 148 * 8) hgrid = 148 * 8;
    if (!hist_ready) os_hist_kernel<<<(unsigned)hgrid, RS_THREADS, 0, st>>>(keys_a, n, n_dev, pass0, npasses, hist);
    // all passes in ONE cooperative kernel: every CTA co-resident, a grid-wide barrier between the passes; at small sizes one
    // CTA per tile, beyond that each CTA takes every grid-th tile
    const int coop_ctas = rs_coop_cta_limit(vals_a != nullptr, rows);
    if (coop_ctas > 0) {
        void *args[] = {&keys_a, &keys_b, &vals_a, &vals_b, &n, &n_dev, (void *)&pass0, (void *)&npasses, &lazy_passes, &hist, &misc, &status};
        const void *fn = vals_a ? rs_all_kernel_for<true>(rows) : rs_all_kernel_for<false>(rows);
        const unsigned grid = (unsigned)std::min<size_t>(ntiles, (size_t)coop_ctas);
        if ((e = cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(RS_THREADS), args, 0, st)) != cudaSuccess) return e;
        if (launches) *launches += (hist_ready ? 0 : 1) + 1;
        return cudaGetLastError();
    }
    // one launch per pass.  High digits first needs both pass counts even: the repair works in place in the first
    // buffer pair, and the conditional full sort must bring the data back there as well
    if (lazy_passes && (((npasses - lazy_passes) & 1) || (npasses & 1))) lazy_passes = 0;
    unsigned long long *kin = keys_a, *kout = keys_b;
    unsigned *vin = vals_a, *vout = vals_b;
    unsigned pass_id = 0;
    const unsigned *run_if = nullptr;
    int nl = hist_ready ? 0 : 1;
    for (int round = 0, first = lazy_passes; round < 2; ++round, first = 0) {
        for (int p = first; p < npasses; ++p, ++pass_id, ++nl) {
            const int shift = 8 * (pass0 + p);
#define RS_LAUNCH(V, R)                                                                                                   \
    os_pass_kernel<V, R><<<(unsigned)ntiles, RS_THREADS, 0, st>>>(kin, vin, n, n_dev, shift, pass_id, hist + p * 256, misc + pass_id, \
                                                                  status, kout, vout, run_if)
            if (vals_a) { if (rows == 2) RS_LAUNCH(true, 2); else if (rows == 4) RS_LAUNCH(true, 4); else if (rows == RS_ROWS_SMALL) RS_LAUNCH(true, RS_ROWS_SMALL); else RS_LAUNCH(true, RS_ROWS_LARGE); }
            else        { if (rows == 2) RS_LAUNCH(false, 2); else if (rows == 4) RS_LAUNCH(false, 4); else if (rows == RS_ROWS_SMALL) RS_LAUNCH(false, RS_ROWS_SMALL); else RS_LAUNCH(false, RS_ROWS_LARGE); }
#undef RS_LAUNCH
            unsigned long long *tk = kin; kin = kout; kout = tk;
            unsigned *tv = vin; vin = vout; vout = tv;
        }
        if (round == 1 || lazy_passes == 0) break;
        size_t rgrid = (n + RS_THREADS - 1) / RS_THREADS;
        if (rgrid > 148 * 8) rgrid = 148 * 8;
        if (vals_a) os_repair_kernel<true><<<(unsigned)rgrid, RS_THREADS, 0, st>>>(keys_a, vals_a, n, n_dev, 8 * (pass0 + lazy_passes), misc + RS_MISC_FLAG);
        else os_repair_kernel<false><<<(unsigned)rgrid, RS_THREADS, 0, st>>>(keys_a, vals_a, n, n_dev, 8 * (pass0 + lazy_passes), misc + RS_MISC_FLAG);
        ++nl;
        run_if = misc + RS_MISC_FLAG;                                       // the passes of the second round run only if a run was too long
    }
    if (launches) *launches += nl;
    if (!lazy_passes && (npasses & 1)) {   // odd number of passes: bring the result back to the first buffer
        if ((e = cudaMemcpyAsync(keys_a, keys_b, n * 8, cudaMemcpyDeviceToDevice, st)) != cudaSuccess) return e;
        if (vals_a && (e = cudaMemcpyAsync(vals_a, vals_b, n * 4, cudaMemcpyDeviceToDevice, st)) != cudaSuccess) return e;
    }
    return cudaGetLastError();
}

// ---- exclusive prefix sum of 32-bit counters, single pass (chained scan with decoupled look-back) ----------------
// Used by the Barnes-Hut build (cells per body -> node offsets).  A CTA takes the next tile of 2048 counters (ticket
// order), publishes the tile's sum, and warp 0 looks back over 32 predecessors at a time until it meets a tile whose
// inclusive prefix is known.  One read and one write of the data.
constexpr int SC_ITEMS = 8, SC_TILE = RS_THREADS * SC_ITEMS;

inline size_t exclusive_scan_temp_bytes(size_t n) { return 64 + ((n + SC_TILE - 1) / SC_TILE) * sizeof(unsigned long long); }

static __global__ void __launch_bounds__(RS_THREADS)
scan_excl_kernel(const unsigned *__restrict__ in, unsigned *__restrict__ out, size_t n, unsigned *__restrict__ ticket,
                 unsigned long long *status)
{
    pdl_enter();
    __shared__ unsigned warp_sums[RS_WARPS];
    __shared__ unsigned s_tile, s_excl;
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const unsigned tile = s_tile;
    const size_t i0 = (size_t)tile * SC_TILE + (size_t)tid * SC_ITEMS;
    unsigned x[SC_ITEMS];
    if (i0 + SC_ITEMS <= n) {                            // in and out come from cudaMalloc: 16-byte aligned rows
        const uint4 a = *reinterpret_cast<const uint4 *>(in + i0), b = *reinterpret_cast<const uint4 *>(in + i0 + 4);
        x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
    } else {
#pragma unroll
        for (int j = 0; j < SC_ITEMS; ++j) x[j] = (i0 + j < n) ? in[i0 + j] : 0u;
    }
    unsigned sum = 0;
#pragma unroll
    for (int j = 0; j < SC_ITEMS; ++j) sum += x[j];
    unsigned total = 0;
    const unsigned before = rs_block_excl_scan(sum, warp_sums, &total);
    if (tid < 32) {                                      // warp 0: publish, look back, publish the inclusive prefix
        if (lane == 0) rs_st_status(status + tile, rs_pack(tile == 0 ? 2u : 1u, total));
        unsigned excl = 0;
        long long t = (long long)tile - 1;
        while (t >= 0) {
            const long long tj = t - lane;
            unsigned long long v = rs_pack(2u, 0u);       // before tile 0: an empty inclusive prefix
            if (tj >= 0) {
                do { v = rs_ld_status(status + tj); } while ((v >> 32) == 0ull);
            }
            const unsigned is_prefix = ((unsigned)(v >> 32) == 2u) ? 1u : 0u;
            const unsigned mask = __ballot_sync(0xffffffffu, is_prefix);
            const int first = mask ? __ffs(mask) - 1 : 31;
            unsigned c = (lane <= first) ? (unsigned)v : 0u;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
            excl += c;
            if (mask) break;
            t -= 32;
        }
        if (lane == 0) {
            s_excl = excl;
            if (tile > 0) rs_st_status(status + tile, rs_pack(2u, excl + total));
        }
    }
    __syncthreads();
    unsigned run = s_excl + before;
    if (i0 + SC_ITEMS <= n) {
        uint4 a, b;
        a.x = run; run += x[0]; a.y = run; run += x[1]; a.z = run; run += x[2]; a.w = run; run += x[3];
        b.x = run; run += x[4]; b.y = run; run += x[5]; b.z = run; run += x[6]; b.w = run;
        *reinterpret_cast<uint4 *>(out + i0) = a;
        *reinterpret_cast<uint4 *>(out + i0 + 4) = b;
    } else {
#pragma unroll
        for (int j = 0; j < SC_ITEMS; ++j) {
            if (i0 + j < n) out[i0 + j] = run;
            run += x[j];
        }
    }
}

// out[i] = in[0] + ... + in[i-1] for i in [0, n).  `temp` holds exclusive_scan_temp_bytes(n).  Asynchronous on `st`.
static inline cudaError_t exclusive_scan_u32(const unsigned *in, unsigned *out, size_t n, void *temp, cudaStream_t st, int *launches = nullptr,
                                              bool temp_zeroed = false)
{
    if (n == 0) return cudaSuccess;
    cudaError_t e = cudaSuccess;
    if (!temp_zeroed && (e = cudaMemsetAsync(temp, 0, exclusive_scan_temp_bytes(n), st)) != cudaSuccess) return e;
    if ((e = launch_pdl(scan_excl_kernel, dim3((unsigned)((n + SC_TILE - 1) / SC_TILE)), dim3(RS_THREADS), 0, st, in, out, n, (unsigned *)temp,
                        (unsigned long long *)((char *)temp + 64))) != cudaSuccess) return e;
    if (launches) *launches += 1;
    return cudaGetLastError();
}

} // namespace nb
