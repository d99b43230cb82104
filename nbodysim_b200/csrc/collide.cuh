// collide.cuh -- phases of the collision pass (see collide.cu for the algorithm), as grid-stride device functions:
// `gtid` of `gthreads` cooperating threads, so that the same code runs as one kernel per phase (large scenes) and inside
// the one-CTA finishing kernel of small scenes (collide.cu).
#pragma once
#include "kernels.h"
#include "radix_sort.cuh"

namespace nb {

constexpr float COL_CELL = 600.0f;   // SpatialGrid::CELL_SIZE, Simulation.hpp:20
// Pair discovery only needs the entries of one grid cell to be findable together, so the (hash, body) entries are
// sorted on the low 16 bits of the hash -- two digit passes instead of four (or eight) -- and the pair loop skips the
// few entries of other cells that share those bits.  Pairs are put in canonical order by their own sort afterwards.
constexpr int COL_GROUP_BITS = 16;
constexpr unsigned long long COL_GROUP_MASK = (1ull << COL_GROUP_BITS) - 1ull;

// (ColArgs, the arguments of a pass, and ColGrid, the screening hash grid, are declared in kernels.h: the fused Barnes-Hut
// walk and the per-step driver handle them too)

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

__device__ __forceinline__ void col_overflow(const ColArgs &a)
{
    atomicExch(&a.counters[2], 1u);
    if (a.status) *reinterpret_cast<volatile unsigned *>(a.status + 1) = 1u;
}

// ---- screening: does ANY pair of bodies sharing a grid cell overlap now? --------------------------------------------------
// resolve() moves nothing unless a pair passes `d.mag_sq() <= r*r` (Simulation.hpp:301), and the first pair to pass it
// in a pass does so on unmoved positions.  So when no two bodies that share a grid cell overlap at the start of the pass,
// the whole pass is the identity -- the common case by far (no collision in the first steps of the shipped scene) -- and
// it is decided without sorting anything: the bodies are inserted into a hash table keyed by their cells (open
// addressing, one linked list of bodies per cell), then every body tests the later-numbered bodies of its cells.  The
// test is resolve()'s own, on every pair sharing a cell: a superset of the sweep pairs, so a miss is impossible and a
// false alarm merely runs the full pass, which applies the reference's rules exactly.

__device__ __forceinline__ bool col_cell_range(const ColBody &b, int &minX, int &maxX, int &minY, int &maxY)
{
    // Simulation.hpp:228-233: AABB = pos -+ radius ; cell = static_cast<int>(coordinate / CELL_SIZE)
    minX = __float2int_rz(__fdiv_rn(__fsub_rn(b.x, b.r), COL_CELL)); maxX = __float2int_rz(__fdiv_rn(__fadd_rn(b.x, b.r), COL_CELL));
    minY = __float2int_rz(__fdiv_rn(__fsub_rn(b.y, b.r), COL_CELL)); maxY = __float2int_rz(__fdiv_rn(__fadd_rn(b.y, b.r), COL_CELL));
    const long long cells = (long long)(maxX - minX + 1) * (long long)(maxY - minY + 1);
    return cells > 0 && cells <= 4096;
}

// A cell's list is split by x STRIPS of COL_STRIP units (the reference's cells are 600 wide and the shipped scene packs 56
// bodies into its busiest cell: a list is a chain of dependent loads, and the pair kernel's time is the longest chain).  A body
// is entered under (cell, strip) for every strip its x interval touches; two bodies whose x intervals overlap -- the only
// ones the sweep pairs up -- both touch the strip in which the overlap STARTS, and that is where the pair is taken, once
// per shared cell as before.  Bodies of different strips cannot form a sweep pair, so nothing is lost.
constexpr float COL_STRIP = 37.5f;
__device__ __forceinline__ int col_strip_of(float x, float strip) { return strip > 0.f ? __float2int_rd(__fdiv_rn(x, strip)) : 0; }
__device__ __forceinline__ unsigned long long col_table_key(int x, int y, int strip)
{
    return ((unsigned long long)((unsigned)strip & 0x7fffffffu) << 33) | (1ull << 32) | (unsigned)col_hash(x, y);
}
// (cells x strips) a body is entered under; false: more than 65,536 (reported as an overflow like a body spanning > 4096 cells)
__device__ __forceinline__ bool col_strip_range(const ColBody &b, float strip, int minX, int maxX, int minY, int maxY, int &s0, int &s1)
{
    s0 = col_strip_of(__fsub_rn(b.x, b.r), strip); s1 = col_strip_of(__fadd_rn(b.x, b.r), strip);
    const long long entries = (long long)(maxX - minX + 1) * (long long)(maxY - minY + 1) * (long long)(s1 - s0 + 1);
    return s1 >= s0 && entries <= 65536;
}

// ---- one (body, cell, strip) unit of the three grid passes --------------------------------------------------------------
__device__ __forceinline__ unsigned col_find_slot(const ColGrid &g, unsigned long long key)
{
    unsigned slot = (unsigned)(key * 0x9E3779B97F4A7C15ull >> 40) & g.tmask;
    for (unsigned probes = 0; probes <= g.tmask; ++probes) {
        const unsigned long long k = g.tkeys[slot];
        if (k == key) return slot;
        if (k == 0ull) break;
        slot = (slot + 1) & g.tmask;
    }
    return 0xffffffffu;
}

struct ColInsertOp {
    __device__ __forceinline__ void operator()(const ColArgs &a, const ColGrid &g, const ColBody &b, unsigned i, int x, int y, int st,
                                               unsigned &) const
    {
        const unsigned long long key = col_table_key(x, y, st);
        unsigned slot = (unsigned)(key * 0x9E3779B97F4A7C15ull >> 40) & g.tmask;
        unsigned probes = 0;
        for (;;) {                                              // find or claim the cell's slot
            const unsigned long long old = atomicCAS(&g.tkeys[slot], 0ull, key);
            if (old == 0ull || old == key) break;
            slot = (slot + 1) & g.tmask;
            if (++probes > g.tmask) break;
        }
        const unsigned e = atomicAdd(&g.flags[1], 1u);
        if (e >= g.ecap || probes > g.tmask) { col_overflow(a); atomicAdd(&g.flags[0], 1u); return; }
        g.edata[e] = make_float4(b.x, b.y, b.r, __uint_as_float(i));
        g.enext[e] = atomicExch(&g.heads[slot], e + 1u);
    }
};

struct ColDetectOp {
    __device__ __forceinline__ void operator()(const ColArgs &, const ColGrid &g, const ColBody &A, unsigned i, int x, int y, int st,
                                               unsigned &overlaps) const
    {
        const unsigned slot = col_find_slot(g, col_table_key(x, y, st));
        if (slot == 0xffffffffu) return;
        for (unsigned e = g.heads[slot]; e != 0u; e = g.enext[e - 1u]) {
            const float4 E = g.edata[e - 1u];
            const unsigned j = __float_as_uint(E.w);
            if (j <= i) continue;                                // every pair at least once (a count; repeats do no harm)
            ColBody B; B.x = E.x; B.y = E.y; B.r = E.z;
            const float dx = __fsub_rn(B.x, A.x), dy = __fsub_rn(B.y, A.y), r = __fadd_rn(A.r, B.r);
            if (!(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) > __fmul_rn(r, r))) ++overlaps;
        }
    }
};

// the three passes share one driver: a body covering one or two (cell, strip) units handles them itself; a body covering
// more -- the shipped scene's central mass of radius 200 sits on 4 cells x 11 strips of the busiest lists -- would be the
// kernel's critical path, so its units are spread over the lanes of its warp
// One body per lane, ALL 32 lanes of the warp calling (valid = false: the lane has no body).
template <typename Op>
__device__ __forceinline__ void col_grid_lane_units(const ColArgs &a, const ColGrid &g, bool valid, const ColBody &b, unsigned i, bool report,
                                                    const Op &op, unsigned &acc)
{
    const unsigned lane = threadIdx.x & 31u;
    int minX = 0, maxX = -1, minY = 0, maxY = -1, s0 = 0, s1 = -1;
    unsigned units = 0;
    if (valid) {
        if (col_cell_range(b, minX, maxX, minY, maxY) && col_strip_range(b, a.strip, minX, maxX, minY, maxY, s0, s1))
            units = (unsigned)((maxX - minX + 1) * (maxY - minY + 1) * (s1 - s0 + 1));
        else if (report) { col_overflow(a); atomicAdd(&g.flags[0], 1u); }          // let the full pass report it
    }
    if (units <= 2u) {
        for (int y = minY; y <= maxY && units; ++y)
            for (int x = minX; x <= maxX; ++x)
                for (int st = s0; st <= s1; ++st) op(a, g, b, i, x, y, st, acc);
    }
    unsigned bigs = __ballot_sync(0xffffffffu, units > 2u);
    while (bigs) {
        const int src = __ffs(bigs) - 1;
        bigs &= bigs - 1u;
        ColBody B;
        B.x = __shfl_sync(0xffffffffu, b.x, src); B.y = __shfl_sync(0xffffffffu, b.y, src); B.r = __shfl_sync(0xffffffffu, b.r, src);
        const unsigned bi = __shfl_sync(0xffffffffu, i, src);
        const int bx0 = __shfl_sync(0xffffffffu, minX, src), bx1 = __shfl_sync(0xffffffffu, maxX, src);
        const int by0 = __shfl_sync(0xffffffffu, minY, src);
        const int bs0 = __shfl_sync(0xffffffffu, s0, src), bs1 = __shfl_sync(0xffffffffu, s1, src);
        const unsigned total = __shfl_sync(0xffffffffu, units, src);
        const unsigned nx = (unsigned)(bx1 - bx0 + 1), ns = (unsigned)(bs1 - bs0 + 1);
        for (unsigned u = lane; u < total; u += 32u) {
            const unsigned st = u % ns, xy = u / ns;
            op(a, g, B, bi, bx0 + (int)(xy % nx), by0 + (int)(xy / nx), bs0 + (int)st, acc);
        }
    }
}

template <typename Op>
__device__ __forceinline__ void col_grid_for_each_unit(const ColArgs &a, const ColGrid &g, unsigned gtid, unsigned gthreads, bool report,
                                                       const Op &op, unsigned &acc)
{
    const unsigned lane = gtid & 31u;
    for (unsigned base = gtid - lane; base < a.n; base += gthreads) {                 // trip count uniform over the warp
        const unsigned i = base + lane;
        ColBody b; b.x = 0.f; b.y = 0.f; b.r = 0.f;
        if (i < a.n) b = col_load(a.posm, a.vel, i);
        col_grid_lane_units(a, g, i < a.n, b, i, report, op, acc);
    }
}

__device__ __forceinline__ void col_grid_insert(const ColArgs &a, const ColGrid &g, unsigned gtid, unsigned gthreads)
{
    unsigned unused = 0;
    col_grid_for_each_unit(a, g, gtid, gthreads, true, ColInsertOp(), unused);
}

__device__ __forceinline__ void col_grid_detect(const ColArgs &a, const ColGrid &g, unsigned gtid, unsigned gthreads)
{
    unsigned overlaps = 0;
    col_grid_for_each_unit(a, g, gtid, gthreads, false, ColDetectOp(), overlaps);
    if (overlaps) atomicAdd(&g.flags[0], overlaps);
}


// ---- union-find (lock-free; a root is always the smallest index of its tree) ------------------------------------
__device__ __forceinline__ unsigned uf_find(unsigned *parent, unsigned x)
{
    unsigned p = __ldcg(parent + x);
    while (p != x) {                                       // path halving; the racing writes are benign (they only
        const unsigned gp = __ldcg(parent + p);            // ever replace a parent by one of its ancestors)
        if (gp != p) __stcg(parent + x, gp);
        x = p;
        p = gp;
    }
    return x;
}
__device__ __forceinline__ void uf_union(unsigned *parent, unsigned a, unsigned b)
{
    for (;;) {
        a = uf_find(parent, a);
        b = uf_find(parent, b);
        if (a == b) return;
        if (a < b) { const unsigned t = a; a = b; b = t; }   // hook the larger root under the smaller
        if (atomicCAS(parent + a, a, b) == a) return;
    }
}

// ---- phases ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void col_phase_init(const ColArgs &a, unsigned gtid, unsigned gthreads)
{
    if (gtid < 8) a.counters[gtid] = 0;
    for (unsigned i = gtid; i < a.n; i += gthreads) { a.parent[i] = i; a.hot[i] = 0; a.hot[a.n + i] = 0; }
}

__device__ __forceinline__ bool col_gate_closed(const ColArgs &a) { return a.gate != nullptr && __ldcg(a.gate) == 0u; }

__device__ __forceinline__ void col_phase_entries(const ColArgs &a, unsigned gtid, unsigned gthreads)
{
    if (col_gate_closed(a)) return;
    for (unsigned i = gtid; i < a.n; i += gthreads) {
        const ColBody b = col_load(a.posm, a.vel, i);
        int minX, maxX, minY, maxY;
        if (!col_cell_range(b, minX, maxX, minY, maxY)) { col_overflow(a); continue; }
        const long long cells = (long long)(maxX - minX + 1) * (long long)(maxY - minY + 1);
        const unsigned base = atomicAdd(&a.counters[0], (unsigned)cells);
        if ((unsigned long long)base + (unsigned long long)cells > a.entry_cap) { col_overflow(a); continue; }
        unsigned k = base;
        for (int y = minY; y <= maxY; ++y)
            for (int x = minX; x <= maxX; ++x) { a.keys_in[k] = col_hash(x, y); a.vals_in[k] = i; ++k; }
    }
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

// Small scenes: the sweep pairs themselves come straight from the hash grid -- body i lists, for every cell it covers,
// the bodies of that cell whose x interval overlaps its own and for which i is the pair's `first` (so every (pair, shared
// cell) is produced exactly once, as by the reference's sweep) -- key (first, second), bit 63 = the pair overlaps now.
// flags[2] counts the pairs, flags[0] the overlapping ones.
constexpr unsigned long long COL_OVERLAP_BIT = 1ull << 63;
struct ColPairsOp {
    __device__ __forceinline__ void operator()(const ColArgs &a, const ColGrid &g, const ColBody &A, unsigned i, int x, int y, int st,
                                               unsigned &overlaps) const
    {
        const unsigned slot = col_find_slot(g, col_table_key(x, y, st));
        if (slot == 0xffffffffu) return;
        const float amin = __fsub_rn(A.x, A.r);
        for (unsigned e = g.heads[slot]; e != 0u; e = g.enext[e - 1u]) {
            const float4 E = g.edata[e - 1u];
            const unsigned j = __float_as_uint(E.w);
            if (j == i) continue;
            ColBody B; B.x = E.x; B.y = E.y; B.r = E.z;
            unsigned first, second;
            if (!col_sweep_pair(A, i, B, j, first, second) || first != i) continue;
            if (col_strip_of(fmaxf(amin, __fsub_rn(B.x, B.r)), a.strip) != st) continue;      // taken in the strip where the overlap starts
            const float dx = __fsub_rn(B.x, A.x), dy = __fsub_rn(B.y, A.y), r = __fadd_rn(A.r, B.r);
            const bool overlap = !(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) > __fmul_rn(r, r));
            const unsigned p = atomicAdd(&g.flags[2], 1u);
            if (p < a.pair_cap) a.pairs[p] = ((unsigned long long)first << a.idx_bits) | second | (overlap ? COL_OVERLAP_BIT : 0ull);
            else col_overflow(a);
            overlaps += overlap ? 1u : 0u;
        }
    }
};
__device__ __forceinline__ void col_grid_pairs(const ColArgs &a, const ColGrid &g, unsigned gtid, unsigned gthreads)
{
    unsigned overlaps = 0;
    col_grid_for_each_unit(a, g, gtid, gthreads, false, ColPairsOp(), overlaps);
    if (overlaps) atomicAdd(&g.flags[0], overlaps);
}

// Enumerate the sweep pairs (one thread per cell entry: the pairs it forms with the later entries of its cell).
//   MODE 0  detect: mark the bodies of pairs that overlap now (Simulation.hpp:301: d.mag_sq() <= r*r) and count them
//   MODE 1  connect: union the two bodies of every sweep pair            (only when something overlaps)
//   MODE 2  emit: the pairs whose component is hot, keyed (component, first, second)
template <int MODE>
__device__ __forceinline__ void col_phase_pairs(const ColArgs &a, unsigned gtid, unsigned gthreads)
{
    if (col_gate_closed(a)) return;
    if (__ldcg(a.counters + 2)) return;                   // overflow: the pass is abandoned (reported through status / sync)
    if (MODE > 0 && __ldcg(a.counters + 4) == 0) return;  // nothing overlaps: no resolve can pass its test
    const unsigned ne = __ldcg(a.counters + 0);
    unsigned overlaps = 0;
    for (unsigned e = gtid; e < ne; e += gthreads) {
        const unsigned long long key = a.keys[e];
        const unsigned ia = a.vals[e];
        const ColBody A = col_load(a.posm, a.vel, ia);
        // the entries are sorted (grouped) by the low COL_GROUP_BITS bits of the hash only: walk the group, keep the same cell
        for (unsigned f = e + 1; f < ne; ++f) {
            const unsigned long long kf = a.keys[f];
            if (((kf ^ key) & COL_GROUP_MASK) != 0ull) break;
            if (kf != key) continue;
            const unsigned ib = a.vals[f];
            if (ib == ia) continue;
            const ColBody B = col_load(a.posm, a.vel, ib);
            unsigned first, second;
            if (!col_sweep_pair(A, ia, B, ib, first, second)) continue;
            if (MODE == 0) {
                const float dx = __fsub_rn(B.x, A.x), dy = __fsub_rn(B.y, A.y), r = __fadd_rn(A.r, B.r);
                if (!(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) > __fmul_rn(r, r))) { a.hot[ia] = 1; a.hot[ib] = 1; ++overlaps; }
            } else if (MODE == 1) {
                uf_union(a.parent, first, second);
            } else {
                const unsigned root = uf_find(a.parent, first);
                if (a.hot[a.n + root]) {
                    const unsigned p = atomicAdd(&a.counters[1], 1u);
                    if (p < a.pair_cap) {
                        unsigned long long k = ((unsigned long long)first << a.idx_bits) | second;      // sorts as (first, second)
                        if (a.rooted) k |= (unsigned long long)root << (2 * a.idx_bits);                // ... inside its component
                        a.pairs[p] = k;
                    } else col_overflow(a);
                }
            }
        }
    }
    if (MODE == 0 && overlaps) atomicAdd(&a.counters[4], overlaps);
}

// a component is hot when one of its bodies is in a pair that overlaps now
__device__ __forceinline__ void col_phase_mark(const ColArgs &a, unsigned gtid, unsigned gthreads)
{
    if (col_gate_closed(a) || __ldcg(a.counters + 2) || __ldcg(a.counters + 4) == 0) return;
    for (unsigned i = gtid; i < a.n; i += gthreads)
        if (a.hot[i]) a.hot[a.n + uf_find(a.parent, i)] = 1;
}

// Simulation::resolve, Simulation.hpp:293-346, unfused IEEE operations in the reference's order.
__device__ __forceinline__ bool col_resolve(float *posm, float *vel, unsigned i, unsigned j)
{
    const size_t gi = blk_index(i, 0), gj = blk_index(j, 0);
    float p1x = posm[gi], p1y = posm[gi + BLK], p2x = posm[gj], p2y = posm[gj + BLK];
    const float dx = __fsub_rn(p2x, p1x), dy = __fsub_rn(p2y, p1y);
    const float r = __fadd_rn(vel[gi + 3 * BLK], vel[gj + 3 * BLK]);
    const float d_sq = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
    const float r_sq = __fmul_rn(r, r);
    if (d_sq > r_sq) return false;
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
        return true;
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
    return true;
}

// One thread per component: the thread whose pair starts a run of equal component labels walks the run in
// (first, second) order.  Components share no body, so runs are independent; inside a run every resolve sees
// what the previous one wrote (one thread, program order).
__device__ __forceinline__ void col_phase_resolve(const ColArgs &a, const unsigned long long *__restrict__ pairs_sorted,
                                                  unsigned gtid, unsigned gthreads)
{
    if (col_gate_closed(a) || __ldcg(a.counters + 2)) return;
    const unsigned np = min(__ldcg(a.counters + 1), a.pair_cap);
    const int b = a.idx_bits, rs = a.rooted ? 2 * b : 63;
    const unsigned long long imask = (1ull << b) - 1ull;
    unsigned resolved = 0;
    for (unsigned p = gtid; p < np; p += gthreads) {
        const unsigned long long k0 = pairs_sorted[p];
        const unsigned long long root = a.rooted ? (k0 >> rs) : 0ull;
        if (p > 0 && (a.rooted ? (pairs_sorted[p - 1] >> rs) : 0ull) == root) continue;      // not the start of a run
        unsigned q = p;
        unsigned long long k = k0;
        for (;;) {
            if (col_resolve(a.posm, a.vel, (unsigned)((k >> b) & imask), (unsigned)(k & imask))) ++resolved;
            if (++q >= np) break;
            k = pairs_sorted[q];
            if ((a.rooted ? (k >> rs) : 0ull) != root) break;
        }
    }
    if (resolved) atomicAdd(&a.counters[3], resolved);
}

} // namespace nb
