// collide.cu -- the reference's collision pass, Simulation::collide + resolve (Simulation.hpp:18-47,
// 216-346), on the GPU (2-D fp32).  SURVEY.md section 8f rank 3.
//
// Broad phase as the reference: hash grid of 600-unit cells over each body's AABB (int-truncated cell
// range, SpatialGrid::hash_position), inside a cell a sweep on x -- two bodies pair up when their x
// intervals overlap, ordered (the one whose interval starts first, the other), once per shared cell.
// The reference resolves pairs in std::unordered_map iteration order, which is unspecified (SURVEY.md
// section 2); here, as in oracle/nbody_oracle.c, the canonical order is "sorted by (first, second)".
// Whenever the colliding pairs of a step are pairwise disjoint the outcome is order-independent and
// equals the reference's bit for bit; chains are resolved deterministically in the canonical order.
//
// Parallel: cell entries, sort, pair discovery, the "hot body" filter (a pair that does not overlap
// now can only start to overlap if an earlier resolve moves one of its bodies, i.e. if one of them is
// in some overlapping pair -- everything else is dropped without changing the result), pair sort.
// Serial: resolve() itself, because each call may read what the previous one wrote (as in the
// reference); collisions are rare events (none in the first steps of the shipped scene).
#include "kernels.h"
#include <cub/device/device_radix_sort.cuh>
#include "radix_sort.cuh"

namespace nb {

constexpr float COL_CELL = 600.0f;   // SpatialGrid::CELL_SIZE, Simulation.hpp:20

__device__ __forceinline__ unsigned long long col_hash(int x, int y)
{
    const unsigned h = (((unsigned)x * 92837111u) ^ ((unsigned)y * 689287499u)) * 15485863u;
    return (unsigned long long)(long long)(int)h;
}

struct ColBody { float x, y, r; };
__device__ __forceinline__ ColBody col_load(const float *posm, const float *vel, unsigned i)
{
    const size_t g = blk_index(i, 0);
    ColBody b;
    b.x = posm[g]; b.y = posm[g + BLK]; b.r = vel[g + 3 * BLK];
    return b;
}

// counters: [0] cell entries, [1] candidate pairs, [2] overflow flag, [3] pairs resolved (narrow test passed)
__global__ void __launch_bounds__(256)
col_entries_kernel(const float *__restrict__ posm, const float *__restrict__ vel, unsigned n,
                   unsigned long long *__restrict__ keys, unsigned *__restrict__ vals, unsigned cap,
                   unsigned *__restrict__ counters)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const ColBody b = col_load(posm, vel, i);
    // Simulation.hpp:228-233: AABB = pos -+ radius ; cell = static_cast<int>(coordinate / CELL_SIZE)
    const int minX = __float2int_rz(__fdiv_rn(__fsub_rn(b.x, b.r), COL_CELL)), maxX = __float2int_rz(__fdiv_rn(__fadd_rn(b.x, b.r), COL_CELL));
    const int minY = __float2int_rz(__fdiv_rn(__fsub_rn(b.y, b.r), COL_CELL)), maxY = __float2int_rz(__fdiv_rn(__fadd_rn(b.y, b.r), COL_CELL));
    const long long cells = (long long)(maxX - minX + 1) * (long long)(maxY - minY + 1);
    if (cells <= 0 || cells > 4096) { atomicExch(&counters[2], 1u); return; }
    const unsigned base = atomicAdd(&counters[0], (unsigned)cells);
    if ((unsigned long long)base + (unsigned long long)cells > cap) { atomicExch(&counters[2], 1u); return; }
    unsigned k = base;
    for (int y = minY; y <= maxY; ++y)
        for (int x = minX; x <= maxX; ++x) { keys[k] = col_hash(x, y); vals[k] = i; ++k; }
}

// x-interval overlap exactly as the sweep sees it (a start sorts before an end at the same x), pair
// ordered by interval start, ties by body index (the reference's unstable sort leaves exact ties open)
__device__ __forceinline__ bool col_sweep_pair(const ColBody &a, unsigned ia, const ColBody &b, unsigned ib,
                                               unsigned &first, unsigned &second)
{
    const float amin = __fsub_rn(a.x, a.r), amax = __fadd_rn(a.x, a.r);
    const float bmin = __fsub_rn(b.x, b.r), bmax = __fadd_rn(b.x, b.r);
    if (fmaxf(amin, bmin) > fminf(amax, bmax)) return false;
    const bool a_first = (amin < bmin) || (amin == bmin && ia < ib);
    first = a_first ? ia : ib;
    second = a_first ? ib : ia;
    return true;
}

// pass 0: mark the bodies of pairs that overlap now (Simulation.hpp:301: d.mag_sq() <= r*r)
// pass 1: emit every sweep pair that overlaps now or touches a marked body
__global__ void __launch_bounds__(256)
col_pairs_kernel(const float *__restrict__ posm, const float *__restrict__ vel,
                 const unsigned long long *__restrict__ keys, const unsigned *__restrict__ vals,
                 const unsigned *__restrict__ counters_in, unsigned char *__restrict__ hot, int pass,
                 unsigned long long *__restrict__ pairs, unsigned pair_cap, unsigned *__restrict__ counters, int idx_bits)
{
    const unsigned e = blockIdx.x * blockDim.x + threadIdx.x;
    if (counters_in[2]) return;                            // overflow: the pass is abandoned (nbody_gpu_collide_stats reports it)
    const unsigned ne = counters_in[0];
    if (e >= ne) return;
    const unsigned long long key = keys[e];
    const unsigned ia = vals[e];
    const ColBody a = col_load(posm, vel, ia);
    for (unsigned f = e + 1; f < ne && keys[f] == key; ++f) {
        const unsigned ib = vals[f];
        if (ib == ia) continue;
        const ColBody b = col_load(posm, vel, ib);
        unsigned first, second;
        if (!col_sweep_pair(a, ia, b, ib, first, second)) continue;
        const float dx = __fsub_rn(b.x, a.x), dy = __fsub_rn(b.y, a.y), r = __fadd_rn(a.r, b.r);
        const bool overlap = !(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) > __fmul_rn(r, r));
        if (pass == 0) {
            if (overlap) { hot[ia] = 1; hot[ib] = 1; }
        } else if (overlap || hot[ia] || hot[ib]) {
            const unsigned p = atomicAdd(&counters[1], 1u);
            if (p < pair_cap) pairs[p] = ((unsigned long long)first << idx_bits) | second;   // sorts as (first, second)
            else atomicExch(&counters[2], 1u);
        }
    }
}

// Simulation::resolve, Simulation.hpp:293-346, unfused IEEE operations in the reference's order.
__device__ void col_resolve(float *posm, float *vel, unsigned i, unsigned j, unsigned *resolved)
{
    const size_t gi = blk_index(i, 0), gj = blk_index(j, 0);
    float p1x = posm[gi], p1y = posm[gi + BLK], p2x = posm[gj], p2y = posm[gj + BLK];
    const float dx = __fsub_rn(p2x, p1x), dy = __fsub_rn(p2y, p1y);
    const float r = __fadd_rn(vel[gi + 3 * BLK], vel[gj + 3 * BLK]);
    const float d_sq = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
    const float r_sq = __fmul_rn(r, r);
    if (d_sq > r_sq) return;
    ++*resolved;
    float v1x = vel[gi], v1y = vel[gi + BLK], v2x = vel[gj], v2y = vel[gj + BLK];
    const float vx = __fsub_rn(v2x, v1x), vy = __fsub_rn(v2y, v1y);
    const float d_dot_v = __fadd_rn(__fmul_rn(dx, vx), __fmul_rn(dy, vy));
    const float m1 = posm[gi + 3 * BLK], m2 = posm[gj + 3 * BLK];
    const float msum = __fadd_rn(m1, m2);
    const float w1 = __fdiv_rn(m2, msum), w2 = __fdiv_rn(m1, msum);
    if (d_dot_v >= 0.0f && !(dx == 0.0f && dy == 0.0f)) {
        const float k = __fsub_rn(__fdiv_rn(r, __fsqrt_rn(d_sq)), 1.0f);
        const float tx = __fmul_rn(dx, k), ty = __fmul_rn(dy, k);
        posm[gi] = __fsub_rn(p1x, __fmul_rn(tx, w1)); posm[gi + BLK] = __fsub_rn(p1y, __fmul_rn(ty, w1));
        posm[gj] = __fadd_rn(p2x, __fmul_rn(tx, w2)); posm[gj + BLK] = __fadd_rn(p2y, __fmul_rn(ty, w2));
        return;
    }
    const float v_sq = __fadd_rn(__fmul_rn(vx, vx), __fmul_rn(vy, vy));
    float disc = __fsub_rn(__fmul_rn(d_dot_v, d_dot_v), __fmul_rn(v_sq, __fsub_rn(d_sq, r_sq)));
    if (disc < 0.0f) disc = 0.0f;
    const float t = __fdiv_rn(__fadd_rn(d_dot_v, __fsqrt_rn(disc)), v_sq);
    p1x = __fsub_rn(p1x, __fmul_rn(v1x, t)); p1y = __fsub_rn(p1y, __fmul_rn(v1y, t));
    p2x = __fsub_rn(p2x, __fmul_rn(v2x, t)); p2y = __fsub_rn(p2y, __fmul_rn(v2y, t));
    const float ndx = __fsub_rn(p2x, p1x), ndy = __fsub_rn(p2y, p1y);
    const float nd_dot_v = __fadd_rn(__fmul_rn(ndx, vx), __fmul_rn(ndy, vy));
    const float nd_sq = __fadd_rn(__fmul_rn(ndx, ndx), __fmul_rn(ndy, ndy));
    const float k = __fdiv_rn(__fmul_rn(1.5f, nd_dot_v), nd_sq);
    const float tx = __fmul_rn(ndx, k), ty = __fmul_rn(ndy, k);
    const float n1x = __fadd_rn(v1x, __fmul_rn(tx, w1)), n1y = __fadd_rn(v1y, __fmul_rn(ty, w1));
    const float n2x = __fsub_rn(v2x, __fmul_rn(tx, w2)), n2y = __fsub_rn(v2y, __fmul_rn(ty, w2));
    vel[gi] = n1x; vel[gi + BLK] = n1y; vel[gj] = n2x; vel[gj + BLK] = n2y;
    posm[gi] = __fadd_rn(p1x, __fmul_rn(n1x, t)); posm[gi + BLK] = __fadd_rn(p1y, __fmul_rn(n1y, t));
    posm[gj] = __fadd_rn(p2x, __fmul_rn(n2x, t)); posm[gj + BLK] = __fadd_rn(p2y, __fmul_rn(n2y, t));
}

__global__ void col_resolve_kernel(float *posm, float *vel, const unsigned long long *__restrict__ pairs,
                                   unsigned pair_cap, unsigned *counters, int idx_bits)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (counters[2]) { counters[3] = 0; return; }
    const unsigned np = min(counters[1], pair_cap);
    unsigned resolved = 0;
    for (unsigned p = 0; p < np; ++p)
        col_resolve(posm, vel, (unsigned)(pairs[p] >> idx_bits), (unsigned)(pairs[p] & ((1ull << idx_bits) - 1ull)), &resolved);
    counters[3] = resolved;
}

cudaError_t CollideWorkspace::alloc(size_t n)
{
    cudaError_t e;
    entry_cap = (unsigned)std::min<size_t>(4 * n + 4096, 0x7fffffffu);
    pair_cap = (unsigned)std::min<size_t>(64 * n + 65536, (size_t)16 << 20);   // candidate pairs kept by the hot-body filter
#define COL_ALLOC(p, bytes) if ((e = cudaMalloc((void **)&(p), (bytes))) != cudaSuccess) return e;
    COL_ALLOC(keys_in, (size_t)entry_cap * 8) COL_ALLOC(keys, (size_t)entry_cap * 8)
    COL_ALLOC(vals_in, (size_t)entry_cap * 4) COL_ALLOC(vals, (size_t)entry_cap * 4)
    COL_ALLOC(pairs_in, (size_t)pair_cap * 8) COL_ALLOC(pairs, (size_t)pair_cap * 8)
    COL_ALLOC(hot, n) COL_ALLOC(counters, 16)
    size_t t1 = 0, t2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, t1, (unsigned long long *)nullptr, (unsigned long long *)nullptr,
                                    (unsigned *)nullptr, (unsigned *)nullptr, (int)entry_cap, 0, 64);
    cub::DeviceRadixSort::SortKeys(nullptr, t2, (unsigned long long *)nullptr, (unsigned long long *)nullptr, (int)pair_cap, 0, 64);
    temp_bytes = std::max(std::max(t1, t2), radix_sort_temp_bytes(std::max<size_t>(entry_cap, pair_cap)));
    COL_ALLOC(temp, temp_bytes)
#undef COL_ALLOC
    n_cap = n;
    return cudaSuccess;
}

void CollideWorkspace::release()
{
    void *ptrs[] = {keys_in, keys, vals_in, vals, pairs_in, pairs, hot, counters, temp};
    for (void *p : ptrs) if (p) cudaFree(p);
    *this = CollideWorkspace();
}

// One collision pass over the first n bodies, in place on posm / vel.  Fully asynchronous: the numbers of cell
// entries and candidate pairs stay on the device -- the grids are sized for the buffers' capacities, the kernels
// and sorts read the live counts (CTAs beyond them exit at once) -- so the pass neither stalls the stream for a
// read-back nor prevents a whole Simulation::step() from being captured in a CUDA graph.
cudaError_t CollideWorkspace::run(float *posm, float *vel, size_t n, cudaStream_t st, int *launches)
{
    if (n == 0 || n > n_cap) return cudaErrorInvalidValue;
    cudaError_t e;
    unsigned *cnt = (unsigned *)counters;
    if ((e = cudaMemsetAsync(cnt, 0, 16, st)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(hot, 0, n, st)) != cudaSuccess) return e;
    const unsigned gn = (unsigned)((n + 255) / 256);
    col_entries_kernel<<<gn, 256, 0, st>>>(posm, vel, (unsigned)n, (unsigned long long *)keys_in, (unsigned *)vals_in, entry_cap, cnt);
    if (launches) *launches += 1;
    // Only the GROUPING of equal cell hashes matters to the pair discovery (pairs are put in canonical order by the
    // second sort), and the hashes are sign-extended 32-bit values: sorting their low 32 bits groups them.
    // A pair key packs (first, second) into 2 x idx_bits bits, so small scenes need 4 digit passes instead of 8.
    int idx_bits = 1;
    while (((size_t)1 << idx_bits) < n) ++idx_bits;
    const int pair_bits = std::min(64, ((2 * idx_bits + 7) / 8) * 8);
    size_t tb = temp_bytes;
    if (own_sort) {
        if ((e = radix_sort_u64((unsigned long long *)keys_in, (unsigned long long *)keys, (unsigned *)vals_in, (unsigned *)vals,
                                entry_cap, temp, st, 0, 32, launches, cnt + 0)) != cudaSuccess) return e;
        std::swap(keys_in, keys);
        std::swap(vals_in, vals);
    } else {   // library comparison path: needs the count on the host
        unsigned h[4];
        if ((e = stats(st, h)) != cudaSuccess) return e;
        if (h[2] || h[0] == 0) return cudaSuccess;
        if ((e = cub::DeviceRadixSort::SortPairs(temp, tb, (const unsigned long long *)keys_in, (unsigned long long *)keys,
                                                 (const unsigned *)vals_in, (unsigned *)vals, (int)h[0], 0, 64, st)) != cudaSuccess) return e;
    }
    const unsigned ge = (entry_cap + 255) / 256;
    for (int pass = 0; pass < 2; ++pass)
        col_pairs_kernel<<<ge, 256, 0, st>>>(posm, vel, (const unsigned long long *)keys, (const unsigned *)vals, cnt,
                                             (unsigned char *)hot, pass, (unsigned long long *)pairs_in, pair_cap, cnt, idx_bits);
    if (launches) *launches += 2;
    tb = temp_bytes;
    if (own_sort) {
        if ((e = radix_sort_u64((unsigned long long *)pairs_in, (unsigned long long *)pairs, nullptr, nullptr, pair_cap, temp, st,
                                0, pair_bits, launches, cnt + 1)) != cudaSuccess) return e;
        std::swap(pairs_in, pairs);
    } else {
        unsigned h[4];
        if ((e = stats(st, h)) != cudaSuccess) return e;
        if (h[2] || h[1] == 0) return cudaSuccess;
        if ((e = cub::DeviceRadixSort::SortKeys(temp, tb, (const unsigned long long *)pairs_in, (unsigned long long *)pairs,
                                                (int)std::min(h[1], pair_cap), 0, 64, st)) != cudaSuccess) return e;
    }
    col_resolve_kernel<<<1, 32, 0, st>>>(posm, vel, (const unsigned long long *)pairs, pair_cap, cnt, idx_bits);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

// counters of the last pass: candidate pairs, resolved pairs, overflow flag (synchronises)
cudaError_t CollideWorkspace::stats(cudaStream_t st, unsigned out[4])
{
    cudaError_t e;
    if ((e = cudaMemcpyAsync(out, counters, 16, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
    return cudaStreamSynchronize(st);
}

} // namespace nb
