// barnes_hut.cu -- the reference's shipped force algorithm, rebuilt for the GPU (2-D, like the reference).
//
// Restates Quadtree::build / insert / propagate (Quadtree.hpp:28-93,157-170,236-258), Quad
// (Quad.hpp:31-57) and the stackless walk Quadtree::acc (Quadtree.hpp:113-155) so that, in refcompat
// arithmetic, accelerations equal the reference's BIT FOR BIT -- including its quirks: one body per
// leaf, COM-distance opening test `size^2 < d^2 theta^2`, far ancestors that contain the target keep
// the target's own mass, and NEAR LEAVES CONTRIBUTE NOTHING (insert() leaves every body Range empty,
// so the leaf loop :133-144 never runs).  `fix_near_leaves` adds the leaf body for near leaves.
//
// Why a parallel build can be bit-faithful to a serial pointer-chasing insert: the reference's tree,
// as a SET of cells, does not depend on insertion order -- a cell exists iff its parent holds >= 2
// distinct positions, child quads come from the fp32 recursion `center + (+-0.5) * size/2`, leaf data
// is the body itself and branch data is sum(child.pos * child.mass) / sum(child.mass) over children
// in quadrant order.  Only the node numbering depends on insertion order, and acc() visits nodes in
// depth-first quadrant order regardless of numbering.  So:
//   1. bounding box -> root quad                                       (Quad::new_containing)
//   2. per body, the quadrant path by the SAME fp32 recursion -> 64-bit key, 2 bits per level,
//      32 levels (Z-order == find_quadrant bit order, Quad.hpp:47-49)
//   3. stable radix sort of (key, body)                                (radix_sort.cuh, hand-written)
//   4. cells owned by each sorted body = the path cells that first appear with it; exclusive scan
//      -> depth-first pre-order node array WITHOUT the reference's empty leaves (they contribute +-0)
//   5. skip pointers (`next`) by binary search on the sorted keys; first child = index + 1
//   6. centres of mass bottom-up in one launch: every leaf climbs towards the root, the LAST child to arrive at a
//      cell (atomic arrival counter) sums the cell's children in quadrant order
//   7. walk: one thread per target (targets in Z-order for coherence), node records from L2.
// Limits: bodies whose positions agree in all 32 levels are merged like the reference's coincident
// bodies (`pos == existing_pos`, masses added in index order); the reference would subdivide
// further if their positions differ beyond that depth.
#include "force_f32_fast.cuh"      // integrate_body_f32
#include "radix_sort.cuh"
#include "cluster_prims.cuh"
#include "collide.cuh"              // the fused walk enters the new positions into the collision pass's screening grid
#include <cstring>

namespace nb {

// DIMS = 2 is the reference's quadtree (2 bits x 32 levels).  DIMS = 3 is the same construction one
// dimension up -- an octree, 3 bits x 21 levels, children ordered (z>cz)<<2 | (y>cy)<<1 | (x>cx) -- for the
// 3-D runs the reference does not have; every rule (box, child centres, one body per leaf, COM sums in child
// order, opening test, near leaves) is the 2-D rule with the z terms appended.
template <int DIMS> struct BhT {
    static constexpr int BITS = DIMS;
    static constexpr int LEVELS = 64 / DIMS;              // 32 or 21
    static constexpr int ALIGN = 64 - LEVELS * BITS;      // keys are left-aligned: 0 or 1 spare low bits
    static constexpr unsigned NCHILD = 1u << DIMS;
};

struct BhRoot { float cx, cy, cz, size; };

// ---- 1. bounding box -------------------------------------------------------------------------
// floats map to unsigned keys whose integer order equals the float order, so min/max reduce with integer
// atomics.  box[0..2] hold the minima x,y,z as the COMPLEMENT of that key and box[3..5] the maxima, so that
// both reduce with atomicMax and "no body seen yet" is 0 for all six words: the per-step reset is part of the
// one memset that also clears the sort / scan scratch and the arrival counters.
__device__ __forceinline__ unsigned f2ord(float f)
{
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned k)
{
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

template <int DIMS>
__global__ void __launch_bounds__(256) bh_bbox_kernel(const float *__restrict__ posm, size_t n, unsigned *box)
{
    float mn[3] = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f};
    float mx[3] = {-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f};
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t g = blk_index(i, 0);
#pragma unroll
        for (int c = 0; c < DIMS; ++c) {
            const float v = posm[g + c * BLK];
            mn[c] = fminf(mn[c], v); mx[c] = fmaxf(mx[c], v);
        }
    }
    // warp, then CTA (shared-memory atomics), then one global atomic per CTA and word: the six words are the same for
    // everybody, and same-address atomics queue up in L2
    __shared__ unsigned sbox[6];
    if (threadIdx.x < 6) sbox[threadIdx.x] = 0u;
    __syncthreads();
#pragma unroll
    for (int c = 0; c < DIMS; ++c) {
        for (int o = 16; o > 0; o >>= 1) {
            mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
            mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
        }
        if ((threadIdx.x & 31) == 0) { atomicMax(&sbox[c], ~f2ord(mn[c])); atomicMax(&sbox[3 + c], f2ord(mx[c])); }
    }
    __syncthreads();
    if (threadIdx.x < 6 && (threadIdx.x % 3) < DIMS) atomicMax(&box[threadIdx.x], sbox[threadIdx.x]);
}

// Quad::new_containing, Quad.hpp:40-44: center = (min+max)*0.5f ; size = max over the axes of the extent.
// Evaluated by every thread of the keys kernel (six cached words, a handful of operations) instead of a
// one-thread kernel of its own.
template <int DIMS>
__device__ __forceinline__ BhRoot bh_root_from_box(const unsigned *__restrict__ box)
{
    float c[3] = {0.f, 0.f, 0.f}, size = 0.f;
#pragma unroll
    for (int a = 0; a < DIMS; ++a) {
        const float lo = ord2f(~box[a]), hi = ord2f(box[3 + a]);
        c[a] = __fmul_rn(__fadd_rn(lo, hi), 0.5f);
        const float ext = __fsub_rn(hi, lo);
        size = (a == 0) ? ext : fmaxf(size, ext);
    }
    BhRoot r;
    r.cx = c[0]; r.cy = c[1]; r.cz = c[2]; r.size = size;
    return r;
}

// Quad::find_quadrant (Quad.hpp:47-49) and Quad::into_quadrant (Quad.hpp:51-57), one level down.
template <int DIMS>
__device__ __forceinline__ unsigned bh_descend(float x, float y, float z, float &cx, float &cy, float &cz, float &size)
{
    unsigned q = ((unsigned)(y > cy) << 1) | (unsigned)(x > cx);
    if (DIMS == 3) q |= (unsigned)(z > cz) << 2;
    const float ns = __fmul_rn(size, 0.5f);
    cx = __fadd_rn(cx, __fmul_rn((q & 1u) ? 0.5f : -0.5f, ns));
    cy = __fadd_rn(cy, __fmul_rn((q & 2u) ? 0.5f : -0.5f, ns));
    if (DIMS == 3) cz = __fadd_rn(cz, __fmul_rn((q & 4u) ? 0.5f : -0.5f, ns));
    size = ns;
    return q;
}

// ---- 2. quadrant-path keys -----------------------------------------------------------------------
// Also: the root quad (kept for the emit kernel) and, with FOLD_HIST, the digit histograms of the radix sort that
// follows -- the keys are counted where they are produced instead of being read again by a histogram kernel.
template <int DIMS, bool FOLD_HIST>
__global__ void __launch_bounds__(256)
bh_keys_kernel(const float *__restrict__ posm, size_t n, const unsigned *__restrict__ box, BhRoot *__restrict__ root_out,
               unsigned long long *__restrict__ keys, unsigned *__restrict__ idx, unsigned *__restrict__ sort_hist)
{
    pdl_enter();
    __shared__ unsigned sh[FOLD_HIST ? RS_MAX_PASSES : 1][256];
    if (FOLD_HIST) rs_hist_clear(sh);
    const BhRoot root = bh_root_from_box<DIMS>(box);
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *root_out = root;
    if (i < n) {
        const size_t g = blk_index(i, 0);
        const float x = posm[g], y = posm[g + BLK], z = (DIMS == 3) ? posm[g + 2 * BLK] : 0.f;
        float cx = root.cx, cy = root.cy, cz = root.cz, size = root.size;
        unsigned long long k = 0;
#pragma unroll 4
        for (int l = 0; l < BhT<DIMS>::LEVELS; ++l) k = (k << BhT<DIMS>::BITS) | bh_descend<DIMS>(x, y, z, cx, cy, cz, size);
        k <<= BhT<DIMS>::ALIGN;
        keys[i] = k;
        idx[i] = (unsigned)i;
        if (FOLD_HIST) rs_hist_add_key(sh, k);
    }
    if (FOLD_HIST) rs_hist_flush(sh, sort_hist);
}

template <int DIMS>
__device__ __forceinline__ int lcp_levels(unsigned long long a, unsigned long long b)
{
    return a == b ? BhT<DIMS>::LEVELS : (__clzll((long long)(a ^ b)) / BhT<DIMS>::BITS);
}

// ---- 4. cells owned by each sorted body ------------------------------------------------------------
// first[s]..leaf[s] are the depths of the cells that first appear with sorted body s; duplicates
// (identical key as the previous body) own nothing.
template <int DIMS>
__global__ void __launch_bounds__(256)
bh_count_kernel(const unsigned long long *__restrict__ keys, size_t n, unsigned *__restrict__ count,
                unsigned char *__restrict__ first, unsigned char *__restrict__ leaf)
{
    pdl_enter();
    const size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    if (s == 0) count[n] = 0;                                // terminates the exclusive scan over count[0..n]
    const unsigned long long k = keys[s];
    if (s > 0 && keys[s - 1] == k) { count[s] = 0; first[s] = 0; leaf[s] = 0; return; }
    const int lp = (s > 0) ? lcp_levels<DIMS>(keys[s - 1], k) : -1;
    size_t t = s + 1;
    while (t < n && keys[t] == k) ++t;                       // skip the run of coincident bodies
    const int ln = (t < n) ? lcp_levels<DIMS>(k, keys[t]) : -1;
    const int leafd = (lp < 0 && ln < 0) ? 0 : max(lp, ln) + 1;   // a lone body is the root leaf
    const int firstd = lp + 1;                                // s == 0 -> 0: owns the root
    count[s] = (unsigned)(leafd - firstd + 1);
    first[s] = (unsigned char)firstd;
    leaf[s] = (unsigned char)leafd;
}

// node record: com/body position, mass, size^2 ; next (0 = end of walk) ; depth | leaf flag ; parent.
// ONE 32-byte record per node, 32-byte aligned, so that a visit touches one sector (two 16-byte vector loads):
//   data = (x, y, mass, size*size)          aux = (z bits [octree], next, depth | leaf << 8 | quadrant << 9, parent)
struct BhNodes {
    float4 *rec;         // 2 x 16 bytes per node
    float4 *quad;        // cx, cy, size, cz (diagnostics / parity tests)
    float4 *slots;       // COM pass: (x, y, z, mass) of each completed child, [parent][quadrant]
    unsigned *slot_cells; // COM pass: cells in that child's subtree (-> skip pointers), [parent][quadrant]
    __device__ __forceinline__ float4 *data(unsigned c) const { return rec + 2 * (size_t)c; }
    __device__ __forceinline__ uint4 *aux(unsigned c) const { return reinterpret_cast<uint4 *>(rec + 2 * (size_t)c + 1); }
};

// First sorted body whose key has the prefix `pp` above bit `shp`, given that body `s` has it: the owner of the cell with
// that prefix.  The bodies with the prefix are a run ending at or after s, so this is the first j in [0, s] with
// (keys[j] >> shp) >= pp.  A search of DEPENDENT loads is what a thread of the build waits for longest, so every round
// issues up to 7 independent probes: one round brackets the answer between two powers of 8 behind s, the following rounds
// cut the bracket in 8 -- about 1 + log8(distance) round trips instead of 2 log2(distance).
__device__ __forceinline__ size_t bh_first_with_prefix(const unsigned long long *__restrict__ keys, size_t s, int shp, unsigned long long pp)
{
    // bracket: lo = a position known to lie before the run (or 0), hi = a position known to lie in the run
    size_t lo = 0, hi = s;
    {
        bool in_run[7];
        size_t pos[7];
#pragma unroll
        for (int j = 0; j < 7; ++j) {
            const size_t back = (size_t)1 << (3 * j);                         // 1, 8, 64, ... 262144 behind s
            pos[j] = s >= back ? s - back : 0;
            in_run[j] = (keys[pos[j]] >> shp) >= pp;
        }
        bool found = false;
#pragma unroll
        for (int j = 0; j < 7; ++j) {
            if (found) continue;
            if (in_run[j]) hi = pos[j];
            else { lo = pos[j] + 1; found = true; }
        }
        // (beyond 8^6 bodies behind s the bracket is [0, hi): the loop below still narrows it 8-fold per round)
    }
    while (hi > lo) {                                                         // invariant: the answer is in [lo, hi]
        const size_t len = hi - lo;
        if (len <= 7) {
            bool in_run[7];
#pragma unroll
            for (int j = 0; j < 7; ++j) in_run[j] = (size_t)j < len ? (keys[lo + j] >> shp) >= pp : true;
            size_t ans = hi;
#pragma unroll
            for (int j = 6; j >= 0; --j) if ((size_t)j < len && in_run[j]) ans = lo + j;
            return ans;
        }
        bool in_run[7];
        size_t pos[7];
#pragma unroll
        for (int j = 0; j < 7; ++j) { pos[j] = lo + (len * (size_t)(j + 1)) / 8; in_run[j] = (keys[pos[j]] >> shp) >= pp; }
        size_t nlo = lo, nhi = hi;
        bool closed = false;
#pragma unroll
        for (int j = 0; j < 7; ++j) {
            if (closed) continue;
            if (in_run[j]) { nhi = pos[j]; closed = true; }
            else nlo = pos[j] + 1;
        }
        lo = nlo; hi = nhi;
    }
    return lo;
}

// ---- 5. emit the pre-order node array -------------------------------------------------------------
template <int DIMS>
__global__ void __launch_bounds__(128)
bh_emit_kernel(const float *__restrict__ posm, const unsigned long long *__restrict__ keys,
               const unsigned *__restrict__ idx, size_t n, const BhRoot *__restrict__ root,
               const unsigned *__restrict__ offs, const unsigned *__restrict__ count,
               const unsigned char *__restrict__ first, const unsigned char *__restrict__ leaf,
               BhNodes nodes, unsigned *__restrict__ arrive, unsigned cap, unsigned *status)
{
    constexpr int BITS = BhT<DIMS>::BITS;
    const size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    // more cells than reserved: the tree is truncated, so raise the context's sticky (host-visible) status word
    if (s == 0 && status) {
        const unsigned cells = offs[n];
        reinterpret_cast<volatile unsigned *>(status)[4] = cells;          // latest cell count: the host grows the reservation early
        if (cells > cap) *reinterpret_cast<volatile unsigned *>(status) = 1u;
    }
    if (s >= n || count[s] == 0) return;
    const unsigned long long k = keys[s];
    const unsigned body = idx[s];
    const size_t g = blk_index(body, 0);
    const float x = posm[g], y = posm[g + BLK], z = (DIMS == 3) ? posm[g + 2 * BLK] : 0.f;
    // merged mass of coincident bodies, added in body-index order (stable sort) like insert() :56-60
    float mass = posm[g + 3 * BLK];
    for (size_t t = s + 1; t < n && keys[t] == k; ++t) mass = __fadd_rn(mass, posm[blk_index(idx[t], 0) + 3 * BLK]);
    const int firstd = first[s], leafd = leaf[s];
    const unsigned off = offs[s];
    float cx = root->cx, cy = root->cy, cz = root->cz, size = root->size;
    for (int d = 0; d <= leafd; ++d) {
        if (d >= firstd) {
            const unsigned c = off + (unsigned)(d - firstd);
            if (c < cap) {
                const bool is_leaf = (d == leafd);
                *nodes.data(c) = make_float4(is_leaf ? x : 0.f, is_leaf ? y : 0.f, is_leaf ? mass : 0.f, __fmul_rn(size, size));
                nodes.quad[c] = make_float4(cx, cy, size, cz);
                // the skip pointer (Node::next) = this cell's index + the number of cells in its subtree: filled in by the
                // centre-of-mass pass, which accumulates subtree sizes on its way up -- no search over the sorted keys
                const unsigned nx = 0;
                // parent cell: the previous cell of this body's chain, or -- for the first owned cell -- the
                // depth d-1 cell of the first sorted body that shares the (d-1)-prefix
                unsigned par = 0xffffffffu;
                if (d > firstd) par = c - 1;
                else if (d == 1) par = 0u;
                else if (d > 1) {
                    const int shp = 64 - BITS * (d - 1);
                    const unsigned long long pp = k >> shp;
                    // first j in [0, s] with (keys[j] >> shp) >= pp: gallop backwards from s, then bisect
                    size_t lo = 0, hi = s, step = 1;
                    while (lo < hi) {
                        const size_t probe = (hi >= lo + step) ? hi - step : lo;
                        if ((keys[probe] >> shp) < pp) { lo = probe + 1; break; }
                        hi = probe;
                        step <<= 1;
                    }
                    while (lo < hi) {
                        const size_t mid = (lo + hi) >> 1;
                        if ((keys[mid] >> shp) >= pp) hi = mid; else lo = mid + 1;
                    }
                    par = offs[lo] + (unsigned)(d - 1 - (int)first[lo]);
                }
                // quadrant of this cell inside its parent = the key bits of level d
                const unsigned qd = (d > 0) ? ((unsigned)(k >> (64 - BITS * d)) & (BhT<DIMS>::NCHILD - 1u)) : 0u;
                // arrive[parent]: low byte = number of children, bits 16.. = which quadrants are occupied
                if (par != 0xffffffffu && par < cap) atomicAdd(&arrive[par], 1u | (1u << (16 + qd)));
                *nodes.aux(c) = make_uint4(__float_as_uint((DIMS == 3 && is_leaf) ? z : 0.f), nx,
                                           (unsigned)d | (is_leaf ? 256u : 0u) | (qd << 9), par);
            }
        }
        if (d < leafd) {
            const unsigned q = (unsigned)(k >> (64 - BITS * (d + 1))) & (BhT<DIMS>::NCHILD - 1u);
            const float ns = __fmul_rn(size, 0.5f);
            cx = __fadd_rn(cx, __fmul_rn((q & 1u) ? 0.5f : -0.5f, ns));
            cy = __fadd_rn(cy, __fmul_rn((q & 2u) ? 0.5f : -0.5f, ns));
            if (DIMS == 3) cz = __fadd_rn(cz, __fmul_rn((q & 4u) ? 0.5f : -0.5f, ns));
            size = ns;
        }
    }
}

// ---- 6. centres of mass (Quadtree::propagate, :236-258) and skip pointers ------------------------------------------------
// One launch, no level barriers: the thread of every leaf climbs towards the root.  A completed node deposits
// (position, mass) and the number of cells of its subtree into its parent's slot for its quadrant, then adds one
// arrival to the parent (bits 8..15 of arrive[]; the low byte holds the number of children and bits 16.. the occupied
// quadrants, both counted by the emit kernel) and stops unless it is the LAST child to arrive.  The last arriver reads
// the occupied slots -- independent loads of one contiguous line, no sibling chasing -- and sums them in quadrant order:
// `pos += child.pos * child.mass; mass += child.mass`, then `pos *= 1/mass` -- exactly the reference's arithmetic
// and order (its empty leaves only ever add +0).  Slots written by other SMs are read with L1-bypassing loads.
// Skip pointers (Node::next, Quadtree.hpp:71-75): in pre-order a subtree is a contiguous run of cells, so
// next[c] = c + cells(subtree of c) (0 after the last cell) -- the subtree sizes ride up with the centres of mass,
// which replaces a gallop-and-bisect search over the sorted keys per cell (the thread of the first body alone did
// ~20 such searches, each up to 25 dependent loads: the emit kernel went 28 -> see profiles/ at n = 25,000).
template <int DIMS>
__global__ void __launch_bounds__(256)
bh_propagate_kernel(BhNodes nodes, size_t n, const unsigned *__restrict__ offs, const unsigned char *__restrict__ first,
                    const unsigned char *__restrict__ leaf, const unsigned *__restrict__ count, unsigned *__restrict__ arrive,
                    unsigned cap)
{
    constexpr unsigned NCHILD = BhT<DIMS>::NCHILD;
    const size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n || count[s] == 0) return;
    const unsigned total = offs[n], m = min(total, cap);
    unsigned c = offs[s] + (unsigned)(leaf[s] - first[s]);           // this body's leaf cell (emitted by the previous launch)
    if (c >= m) return;
    uint4 a = __ldcg(nodes.aux(c));
    const float4 d0 = __ldcg(nodes.data(c));
    float x = d0.x, y = d0.y, z = (DIMS == 3) ? __uint_as_float(a.x) : 0.f, mass = d0.z;
    unsigned cells = 1;                                               // cells in the subtree of c
    reinterpret_cast<unsigned *>(nodes.aux(c))[1] = (c + 1u < total) ? c + 1u : 0u;
    for (;;) {
        const unsigned par = a.w;
        if (par == 0xffffffffu || par >= m) break;                     // reached the root
        const unsigned q = (a.z >> 9) & (NCHILD - 1u);
        // everything about the parent that does not depend on its other children, requested at once
        const unsigned pinfo = __ldcg(&arrive[par]);                   // low byte: number of children (counted by the emit kernel)
        const uint4 pa = __ldcg(nodes.aux(par));
        float px = 0.f, py = 0.f, pz = 0.f, ms = 0.f;
        if ((pinfo & 0xffu) == 1u) {
            // an only child (the chains of single-child cells above two close bodies): nobody to wait for, nothing to
            // publish -- the same sum over "the children in quadrant order", with one term
            px = __fadd_rn(px, __fmul_rn(x, mass));
            py = __fadd_rn(py, __fmul_rn(y, mass));
            if (DIMS == 3) pz = __fadd_rn(pz, __fmul_rn(z, mass));
            ms = __fadd_rn(ms, mass);
            cells += 1;
        } else {
            // the subtree size rides in the slot's unused z word in 2-D, in a parallel array for the octree
            __stcg(&nodes.slots[(size_t)par * NCHILD + q], make_float4(x, y, (DIMS == 3) ? z : __uint_as_float(cells), mass));
            if (DIMS == 3) __stcg(&nodes.slot_cells[(size_t)par * NCHILD + q], cells);
            __threadfence();                                           // my deposit is visible before I announce it
            const unsigned old = atomicAdd(&arrive[par], 0x100u);
            if (((old >> 8) & 0xffu) + 1u != (old & 0xffu)) break;     // a sibling will arrive later and do the work
            __threadfence();
            const unsigned mask = (old >> 16) & 0xffu;
            float4 ch[NCHILD];
            unsigned sub[NCHILD];
#pragma unroll
            for (unsigned k = 0; k < NCHILD; ++k) {
                const bool occ = (mask >> k) & 1u;
                ch[k] = occ ? __ldcg(&nodes.slots[(size_t)par * NCHILD + k]) : make_float4(0.f, 0.f, 0.f, 0.f);
                sub[k] = !occ ? 0u : (DIMS == 3) ? __ldcg(&nodes.slot_cells[(size_t)par * NCHILD + k]) : __float_as_uint(ch[k].z);
            }
            cells = 1;
#pragma unroll
            for (unsigned k = 0; k < NCHILD; ++k) {
                if ((mask >> k) & 1u) {                                // children in quadrant order
                    px = __fadd_rn(px, __fmul_rn(ch[k].x, ch[k].w));
                    py = __fadd_rn(py, __fmul_rn(ch[k].y, ch[k].w));
                    if (DIMS == 3) pz = __fadd_rn(pz, __fmul_rn(ch[k].z, ch[k].w));
                    ms = __fadd_rn(ms, ch[k].w);
                    cells += sub[k];
                }
            }
        }
        if (ms > 0.f) { // Vec2::operator/=: inv = 1/scalar ; x *= inv ; y *= inv
            const float inv = __fdiv_rn(1.0f, ms);
            px = __fmul_rn(px, inv);
            py = __fmul_rn(py, inv);
            if (DIMS == 3) pz = __fmul_rn(pz, inv);
        }
        // position and mass of the parent's record; its fourth word (size^2, written by the emit kernel) stays untouched
        __stcg(reinterpret_cast<float2 *>(nodes.data(par)), make_float2(px, py));
        __stcg(reinterpret_cast<float *>(nodes.data(par)) + 2, ms);
        if (DIMS == 3) __stcg(reinterpret_cast<float *>(nodes.aux(par)), pz);
        reinterpret_cast<unsigned *>(nodes.aux(par))[1] = (par + cells < total) ? par + cells : 0u;
        x = px; y = py; z = pz; mass = ms;
        a = pa;
        c = par;
    }
}

// ---- 5 + 6 in one kernel for everything but the top of the tree -------------------------------------------------------------
// The climb above pays two or three dependent L2 round trips, a fence and an atomic PER TREE LEVEL (22 levels on the
// reference's scene: 44 us of a 200 us step) although almost every cell covers a handful of neighbouring sorted bodies.
// Here a CTA owns BHL_MAIN consecutive sorted bodies and sees a window of BHL_MAIN + BHL_HALO of them.  A cell is LOCAL
// iff its span (owner .. last body) is at most BHL_HALO bodies long -- then the whole subtree lies inside the owner's
// window and is finished there, level by level from the deepest, through shared memory: a thread keeps the datum of the
// chain cell it is working on; when a cell is complete the owner of its parent adds the children up IN QUADRANT ORDER
// (its own chain child first, then the published first cells of the following threads, each of which also says where
// its span ends, i.e. where the next sibling starts) with exactly Quadtree::propagate's arithmetic.  Skip pointers fall
// out: next = first cell of the body after the span.  The halo threads redo the next CTAs' subtrees (no writes), which
// costs nothing next to a round trip through L2.  Only cells LONGER than the halo (a few hundred at n = 25,000) are left
// to the climb: every thread reports the cell where its chain leaves the local part (`gstart`) and counts it as a
// child of its parent; bh_climb_kernel then runs the arrival protocol of bh_propagate_kernel from there.
// The length criterion is one both sides can evaluate: the parent's owner sees its whole window, a child finds the
// owner by the gallop-and-bisect of bh_emit_kernel and the span's end by following its siblings.
template <int DIMS, bool TRACE, int BHL_MAIN, int BHL_HALO>
__global__ void __launch_bounds__(BHL_MAIN + BHL_HALO)
bh_emit_local_kernel(const float *__restrict__ posm, const unsigned long long *__restrict__ keys,
                     const unsigned *__restrict__ idx, size_t n, const BhRoot *__restrict__ root,
                     const unsigned *__restrict__ offs, const unsigned *__restrict__ count,
                     const unsigned char *__restrict__ first, const unsigned char *__restrict__ leaf,
                     BhNodes nodes, unsigned *__restrict__ arrive, unsigned cap, unsigned *status, unsigned *__restrict__ gstart,
                     uint4 *__restrict__ edge_list, unsigned *__restrict__ edge_count, unsigned edge_cap, unsigned long long *trace)
{
    pdl_enter();
    // tuning aid (NBODY_BH_TRACE): per phase, the largest number of SM cycles any thread needed to get there
    const long long trace_t0 = TRACE ? clock64() : 0;
    int trace_k = 0;
#define BHL_TRACE() do { if (TRACE) { const unsigned am = __activemask();                                                          \
                                      const unsigned cyc = __reduce_max_sync(am, (unsigned)(clock64() - trace_t0));                  \
                                      if ((threadIdx.x & 31) == (unsigned)(__ffs(am) - 1)) atomicMax(trace + trace_k, (unsigned long long)cyc); \
                                      ++trace_k; } } while (0)
    constexpr int BITS = BhT<DIMS>::BITS;
    constexpr unsigned NCHILD = BhT<DIMS>::NCHILD;
    constexpr int M = BHL_MAIN, H = BHL_HALO, W = BHL_MAIN + BHL_HALO;
    constexpr unsigned short NOT_LOCAL = 0xffffu;
    __shared__ int sA[W + 1];                 // levels shared with the previous sorted body (-1: none); [W] = the body after the window
    __shared__ unsigned soffs[W + 1];         // first cell of every body of the window
    __shared__ float4 sval[W];                // datum of a thread's FIRST cell once complete: (x, y, z, mass)
    __shared__ unsigned spk[W + 1];           // ... its parent's depth and the window index where its span ends (see below)
    __shared__ unsigned char spulled[W];      // the first cell was added to its parent by the parent's owner in this CTA
    __shared__ int s_maxd;
    const unsigned i = threadIdx.x;
    const size_t s0 = (size_t)blockIdx.x * M, s = s0 + i;
    const unsigned total = offs[n];
    if (s == 0 && status) {                   // more cells than reserved: see bh_emit_kernel
        reinterpret_cast<volatile unsigned *>(status)[4] = total;
        if (total > cap) *reinterpret_cast<volatile unsigned *>(status) = 1u;
    }
    const bool exists = s < n;
    const unsigned long long k = exists ? keys[s] : 0ull;
    sA[i] = (exists && s > 0) ? lcp_levels<DIMS>(keys[s - 1], k) : -1;
    soffs[i] = s <= n ? offs[s] : total;
    if (i == W - 1) {
        const size_t sw = s0 + W;
        sA[W] = sw < n ? lcp_levels<DIMS>(keys[sw - 1], keys[sw]) : -1;
        soffs[W] = sw <= n ? offs[sw] : total;
    }
    spulled[i] = 0;
    if (i == 0) s_maxd = 0;
    const unsigned cnt = exists ? count[s] : 0u;
    const bool main_thread = i < (unsigned)M;
    int firstd = 0, leafd = 0;
    unsigned off = 0, tp = i + 1;            // tp: window index where the span of the cell in `cur` ends
    float cx_ = 0.f, cy_ = 0.f, cz_ = 0.f, cm_ = 0.f;   // `cur`: datum of the deepest unfinished... of the chain cell at depth dcur
    if (cnt) {
        firstd = first[s]; leafd = leaf[s]; off = offs[s];
        const size_t g = blk_index(idx[s], 0);
        cx_ = posm[g]; cy_ = posm[g + BLK]; cz_ = (DIMS == 3) ? posm[g + 2 * BLK] : 0.f;
        // merged mass of coincident bodies, added in body-index order (stable sort) like insert() :56-60
        cm_ = posm[g + 3 * BLK];
        size_t t = s + 1;
        for (; t < n && keys[t] == k; ++t) cm_ = __fadd_rn(cm_, posm[blk_index(idx[t], 0) + 3 * BLK]);
        tp = (unsigned)min(t - s0, (size_t)W + 1);
    }
    __syncthreads();
    BHL_TRACE();   // 0: loads

    // ---- skeleton of the cells this body owns (top-down: the fp32 quad recursion), leaf record complete
    if (cnt && main_thread) {
        float qx = root->cx, qy = root->cy, qz = root->cz, size = root->size;
        for (int d = 0; d <= leafd; ++d) {
            if (d >= firstd) {
                const unsigned c = off + (unsigned)(d - firstd);
                if (c < cap) {
                    const bool is_leaf = (d == leafd);
                    *nodes.data(c) = make_float4(is_leaf ? cx_ : 0.f, is_leaf ? cy_ : 0.f, is_leaf ? cm_ : 0.f, __fmul_rn(size, size));
                    nodes.quad[c] = make_float4(qx, qy, size, qz);
                    const unsigned qd = (d > 0) ? ((unsigned)(k >> (64 - BITS * d)) & (NCHILD - 1u)) : 0u;
                    *nodes.aux(c) = make_uint4(__float_as_uint((DIMS == 3 && is_leaf) ? cz_ : 0.f),
                                               (is_leaf && c + 1u < total) ? c + 1u : 0u,
                                               (unsigned)d | (is_leaf ? 256u : 0u) | (qd << 9), (d > firstd) ? c - 1u : 0xffffffffu);
                }
            }
            if (d < leafd) {
                const unsigned q = (unsigned)(k >> (64 - BITS * (d + 1))) & (NCHILD - 1u);
                const float ns = __fmul_rn(size, 0.5f);
                qx = __fadd_rn(qx, __fmul_rn((q & 1u) ? 0.5f : -0.5f, ns));
                qy = __fadd_rn(qy, __fmul_rn((q & 2u) ? 0.5f : -0.5f, ns));
                if (DIMS == 3) qz = __fadd_rn(qz, __fmul_rn((q & 4u) ? 0.5f : -0.5f, ns));
                size = ns;
            }
        }
    }

    BHL_TRACE();   // 1: skeleton
    // ---- centres of mass of the local cells, deepest level first (one CTA barrier per level)
    // (Tried: letting every thread climb as far as its children are published, rounds instead of levels -- same time: the
    //  loop is bound by the instructions all warps issue per step, not by the barriers.)
    // spk[t] packs what a parent needs to know about thread t's first cell in ONE word: bits 0..7 the levels t shares with its
    // predecessor (= the depth of that cell's parent; 0xff: none), bits 8.. the window index where the cell's span ends
    // (NOT_LOCAL until the cell is complete, and for good if it is not local) -- one shared-memory round trip per child.
    auto span_ok = [&](unsigned t) { return t <= (unsigned)W && t - i <= (unsigned)H; };
    bool alive = cnt != 0 && span_ok(tp);     // false: this chain continues in the climb (or the body owns nothing)
    int dcur = leafd;
    const unsigned myA = (unsigned)sA[i] & 0xffu;
    if (alive && firstd == leafd) { sval[i] = make_float4(cx_, cy_, cz_, cm_); spk[i] = myA | (tp << 8); }
    else spk[i] = myA | ((unsigned)NOT_LOCAL << 8);
    if (i == W - 1) spk[W] = ((unsigned)sA[W] & 0xffu) | ((unsigned)NOT_LOCAL << 8);
    if (cnt) atomicMax(&s_maxd, leafd);
    __syncthreads();
    const long long trace_lv0 = TRACE ? clock64() : 0;
    int trace_iters = 0;
    for (int d = s_maxd - 1; d >= 0; --d) {
        if (TRACE) ++trace_iters;
        if (alive && d >= firstd && d < leafd) {              // the cell at depth d of this chain; `cur` is its first child
            float px = 0.f, py = 0.f, pz = 0.f, ms = 0.f;
            px = __fadd_rn(px, __fmul_rn(cx_, cm_));
            py = __fadd_rn(py, __fmul_rn(cy_, cm_));
            if (DIMS == 3) pz = __fadd_rn(pz, __fmul_rn(cz_, cm_));
            ms = __fadd_rn(ms, cm_);
            unsigned t = tp, kids[NCHILD - 1];
            int nk = 0;
            bool ok = true, more = true;
#pragma unroll
            for (int c = 0; c < (int)NCHILD - 1; ++c) {        // the following children: first cells of later threads, in key order
                if (more) {
                    const unsigned tt = min(t, (unsigned)W);
                    const unsigned w_ = spk[tt];                // requested together with the datum it may announce
                    const float4 ch = sval[min(tt, (unsigned)W - 1u)];
                    more = t <= (unsigned)W && (w_ & 0xffu) == (unsigned)d;
                    if (more) {
                        const unsigned e = w_ >> 8;
                        if (e == (unsigned)NOT_LOCAL) { ok = false; more = false; }
                        else {
                            px = __fadd_rn(px, __fmul_rn(ch.x, ch.w));
                            py = __fadd_rn(py, __fmul_rn(ch.y, ch.w));
                            if (DIMS == 3) pz = __fadd_rn(pz, __fmul_rn(ch.z, ch.w));
                            ms = __fadd_rn(ms, ch.w);
                            kids[c] = t; nk = c + 1;
                            t = e;
                        }
                    }
                }
            }
            if (ok && span_ok(t)) {
#pragma unroll
                for (int c = 0; c < (int)NCHILD - 1; ++c) if (c < nk) spulled[kids[c]] = 1;
                if (ms > 0.f) { // Vec2::operator/=: inv = 1/scalar ; x *= inv ; y *= inv   (1/m correctly rounded: rcp.rn)
                    const float inv = __frcp_rn(ms);
                    px = __fmul_rn(px, inv);
                    py = __fmul_rn(py, inv);
                    if (DIMS == 3) pz = __fmul_rn(pz, inv);
                }
                cx_ = px; cy_ = py; cz_ = pz; cm_ = ms;
                tp = t; dcur = d;
                if (d == firstd) { sval[i] = make_float4(px, py, pz, ms); spk[i] = myA | (t << 8); }
                const unsigned c = off + (unsigned)(d - firstd);
                if (main_thread && c < cap) {
                    *reinterpret_cast<float2 *>(nodes.data(c)) = make_float2(px, py);
                    reinterpret_cast<float *>(nodes.data(c))[2] = ms;
                    if (DIMS == 3) reinterpret_cast<float *>(nodes.aux(c))[0] = pz;
                    reinterpret_cast<unsigned *>(nodes.aux(c))[1] = (s0 + t < n) ? soffs[t] : 0u;
                }
            } else alive = false;
        }
        if (TRACE && i == 0) { atomicAdd(trace + 8 + (blockIdx.x == 0 ? 0 : 1), 1ull); atomicMax(trace + 10, (unsigned long long)(s_maxd - d)); }
        if (!__syncthreads_or(alive && dcur > firstd)) break;
    }
    __syncthreads();
    if (TRACE && i == 0 && blockIdx.x < 1000) { trace[16 + 2 * blockIdx.x] = (unsigned long long)(clock64() - trace_lv0); trace[17 + 2 * blockIdx.x] = (unsigned long long)trace_iters; }

    BHL_TRACE();   // 2: local levels
    // ---- hand-over to the climb: count every cell that the local part did not attach to its parent as a child of that parent
    if (!(cnt && main_thread)) return;
    unsigned start = 0xffffffffu;
    // cells of the chain above the local part (depths firstd .. dcur-1) and the local part's top cell (depth dcur) hang from the chain
    // small scenes also LIST every such (child, parent) edge for the one-CTA climb: (child cell, parent cell, quadrant | depth << 8
    // | "the child's datum is complete" << 16)
    auto list_edge = [&](unsigned child, unsigned parent, unsigned qd, int d, bool complete) {
        if (!edge_list) return;
        const unsigned e = atomicAdd(edge_count, 1u);
        if (e < edge_cap) edge_list[e] = make_uint4(child, parent, qd | ((unsigned)d << 8) | (complete ? 65536u : 0u), 0u);
    };
    for (int d = dcur; d > firstd; --d) {
        const unsigned c = off + (unsigned)(d - firstd), par = c - 1u, qd = (unsigned)(k >> (64 - BITS * d)) & (NCHILD - 1u);
        if (par < cap) atomicAdd(&arrive[par], 1u | (1u << (16 + qd)));
        list_edge(c, par, qd, d, d == dcur);
    }
    if (dcur > firstd) start = off + (unsigned)(dcur - firstd);
    if (firstd > 0 && !(dcur == firstd && spulled[i])) {
        const int shp = 64 - BITS * (firstd - 1);
        const unsigned long long pp = k >> shp;
        // Is the parent local to ITS owner's CTA (then that CTA's halo threads redid this cell and added it there)?  Only if
        // this cell is local and the parent's span -- its owner .. the end of my last sibling -- is at most H bodies long:
        // the body H + 1 places before the span's end must lie outside the parent (ONE key; no search for the owner)
        bool parent_local = false;
        if (dcur == firstd && alive) {
            unsigned t = tp;
            bool ok = true;
            while (t <= (unsigned)W && sA[t] == firstd - 1) {
                const unsigned e = spk[t] >> 8;
                if (e == (unsigned)NOT_LOCAL) { ok = false; break; }
                t = e;
            }
            if (ok && t <= (unsigned)W) {
                const size_t e = s0 + t;
                parent_local = e <= (size_t)H || (keys[e - H - 1] >> shp) != pp;
            }
        }
        if (!parent_local) {
            const size_t lo = bh_first_with_prefix(keys, s, shp, pp);        // owner of the parent cell
            const unsigned par = offs[lo] + (unsigned)(firstd - 1 - (int)first[lo]);
            if (off < cap) reinterpret_cast<unsigned *>(nodes.aux(off))[3] = par;
            const unsigned qd = (unsigned)(k >> (64 - BITS * firstd)) & (NCHILD - 1u);
            if (par < cap) atomicAdd(&arrive[par], 1u | (1u << (16 + qd)));
            list_edge(off, par, qd, firstd, dcur == firstd);
            if (dcur == firstd) start = off;
        }
    }
    gstart[s] = start;
    BHL_TRACE();   // 3: hand-over
#undef BHL_TRACE
}

// The climb over the cells the local kernel left: bh_propagate_kernel's protocol, started from the cell where a chain leaves
// the local part instead of from the leaves.  Returns (cells finished, of which with several children) for the tuning trace.
template <int DIMS>
__device__ __forceinline__ uint2 bh_climb_chain(const BhNodes &nodes, unsigned c, unsigned total, unsigned m, unsigned *__restrict__ arrive)
{
    constexpr unsigned NCHILD = BhT<DIMS>::NCHILD;
    unsigned levels = 0, waits = 0;
    uint4 a = __ldcg(nodes.aux(c));
    const float4 d0 = __ldcg(nodes.data(c));
    float x = d0.x, y = d0.y, z = (DIMS == 3) ? __uint_as_float(a.x) : 0.f, mass = d0.z;
    unsigned cells = (a.y ? a.y : total) - c;                          // cells in the subtree of c (from its skip pointer)
    for (;;) {
        const unsigned par = a.w;
        if (par == 0xffffffffu || par >= m) break;                     // reached the root
        const unsigned q = (a.z >> 9) & (NCHILD - 1u);
        const unsigned pinfo = __ldcg(&arrive[par]);                   // low byte: number of children
        const uint4 pa = __ldcg(nodes.aux(par));
        float px = 0.f, py = 0.f, pz = 0.f, ms = 0.f;
        if ((pinfo & 0xffu) == 1u) {                                   // an only child: nobody to wait for
            px = __fadd_rn(px, __fmul_rn(x, mass));
            py = __fadd_rn(py, __fmul_rn(y, mass));
            if (DIMS == 3) pz = __fadd_rn(pz, __fmul_rn(z, mass));
            ms = __fadd_rn(ms, mass);
            cells += 1;
        } else {
            __stcg(&nodes.slots[(size_t)par * NCHILD + q], make_float4(x, y, (DIMS == 3) ? z : __uint_as_float(cells), mass));
            if (DIMS == 3) __stcg(&nodes.slot_cells[(size_t)par * NCHILD + q], cells);
            __threadfence();                                           // my deposit is visible before I announce it
            const unsigned old = atomicAdd(&arrive[par], 0x100u);
            if (((old >> 8) & 0xffu) + 1u != (old & 0xffu)) break;     // a sibling will arrive later and do the work
            // no second fence: the loads below are issued only after the atomic has returned (the branch above depends on
            // its result and a GPU does not speculate), every sibling fenced its deposit before ITS arrival, and the loads
            // bypass L1 -- one fence per level instead of two on the build's longest dependent chain
            const unsigned mask = (old >> 16) & 0xffu;
            float4 ch[NCHILD];
            unsigned sub[NCHILD];
#pragma unroll
            for (unsigned k = 0; k < NCHILD; ++k) {
                const bool occ = (mask >> k) & 1u;
                ch[k] = occ ? __ldcg(&nodes.slots[(size_t)par * NCHILD + k]) : make_float4(0.f, 0.f, 0.f, 0.f);
                sub[k] = !occ ? 0u : (DIMS == 3) ? __ldcg(&nodes.slot_cells[(size_t)par * NCHILD + k]) : __float_as_uint(ch[k].z);
            }
            cells = 1;
#pragma unroll
            for (unsigned k = 0; k < NCHILD; ++k) {
                if ((mask >> k) & 1u) {                                // children in quadrant order
                    px = __fadd_rn(px, __fmul_rn(ch[k].x, ch[k].w));
                    py = __fadd_rn(py, __fmul_rn(ch[k].y, ch[k].w));
                    if (DIMS == 3) pz = __fadd_rn(pz, __fmul_rn(ch[k].z, ch[k].w));
                    ms = __fadd_rn(ms, ch[k].w);
                    cells += sub[k];
                }
            }
            ++waits;
        }
        if (ms > 0.f) {                                                // 1/m correctly rounded, as the reference's division
            const float inv = __frcp_rn(ms);
            px = __fmul_rn(px, inv);
            py = __fmul_rn(py, inv);
            if (DIMS == 3) pz = __fmul_rn(pz, inv);
        }
        __stcg(reinterpret_cast<float2 *>(nodes.data(par)), make_float2(px, py));
        __stcg(reinterpret_cast<float *>(nodes.data(par)) + 2, ms);
        if (DIMS == 3) __stcg(reinterpret_cast<float *>(nodes.aux(par)), pz);
        reinterpret_cast<unsigned *>(nodes.aux(par))[1] = (par + cells < total) ? par + cells : 0u;
        x = px; y = py; z = pz; mass = ms;
        a = pa;
        c = par;
        ++levels;
    }
    return make_uint2(levels, waits);
}

template <int DIMS, bool TRACE>
__global__ void __launch_bounds__(256)
bh_climb_kernel(BhNodes nodes, size_t n, const unsigned *__restrict__ offs, const unsigned *__restrict__ count,
                const unsigned *__restrict__ gstart, unsigned *__restrict__ arrive, unsigned cap, unsigned long long *trace)
{
    pdl_enter();
    const long long trace_t0 = TRACE ? clock64() : 0;
    const size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n || count[s] == 0) return;
    const unsigned c = gstart[s];
    const unsigned total = offs[n], m = min(total, cap);
    if (c >= m) return;                                                // 0xffffffff: nothing of this chain is left
    const uint2 t = bh_climb_chain<DIMS>(nodes, c, total, m, arrive);
    if (TRACE) {
        atomicAdd(trace + 0, 1ull);                                    // chains that reach the climb
        atomicMax(trace + 1, (unsigned long long)t.x);                 // most cells finished by one thread ...
        atomicMax(trace + 2, (unsigned long long)t.y);                 // ... of which with more than one child (slot + fence + atomic)
        atomicMax(trace + 3, (unsigned long long)(clock64() - trace_t0));
    }
}

// Small scenes: the top of the tree in the SHARED MEMORY of one CTA.  Whatever climbs through global memory pays one or two
// L2 round trips (and a fence) per tree level -- ~3,000 cycles x 8 levels on the shipped scene, whether many CTAs use atomics or
// one CTA uses barriers (both measured, profiles/).  But the cells involved are few: the local kernel lists every (child, parent)
// edge it could not close, ~1,000 at 25,000 bodies.  One CTA loads the list and the data of the complete children once (two
// round trips for everything), finds every edge's parent through a small hash table, and then sums level by level out of shared
// memory -- children in quadrant order, Quadtree::propagate's arithmetic -- storing each finished cell's record on the way.
// More than BTC_MAX_EDGES edges (a pathological scene at these sizes): the CTA runs the atomic protocol over `gstart` instead.
constexpr int BTC_THREADS = 1024, BTC_MAX_EDGES = 2047, BTC_HASH = 4096;
template <int DIMS> struct BtcSmem {
    float4 val[BTC_MAX_EDGES + 1];                        // (x, y, z, mass) of node e = the child cell of edge e; [E] = the root
    unsigned cell[BTC_MAX_EDGES + 1], sub[BTC_MAX_EDGES + 1], info[BTC_MAX_EDGES + 1];   // cell index, cells in its subtree, quadrant | depth << 8 | complete << 16
    unsigned short pid[BTC_MAX_EDGES + 1];                // node of the parent cell
    unsigned short kids[(BTC_MAX_EDGES + 1) * (1 << DIMS)];   // node of the child in each quadrant, 0xffff = none
    unsigned hkey[BTC_HASH];                              // cell index + 1 -> node (open addressing)
    unsigned short hval[BTC_HASH];
    int maxd, bad;
};

template <int DIMS, bool TRACE>
__global__ void __launch_bounds__(BTC_THREADS, 1)
bh_top_cta_kernel(BhNodes nodes, size_t n, const unsigned *__restrict__ offs, const unsigned *__restrict__ count, const uint4 *__restrict__ edges,
                  const unsigned *__restrict__ nedges, unsigned edge_cap, const unsigned *__restrict__ gstart, unsigned *__restrict__ arrive,
                  unsigned cap, unsigned edge_limit, unsigned long long *trace)
{
    pdl_enter();
    constexpr unsigned NCHILD = BhT<DIMS>::NCHILD;
    extern __shared__ __align__(16) unsigned char btc_raw[];
    BtcSmem<DIMS> &sm = *reinterpret_cast<BtcSmem<DIMS> *>(btc_raw);
    const long long trace_t0 = TRACE ? clock64() : 0;
    const unsigned tid = threadIdx.x;
    const unsigned total = offs[n], m = min(total, cap), E = *nedges;
    if (E == 0u) return;                                               // everything was local (n <= halo)
    if (E > edge_limit || E > edge_cap || total > cap) {               // too many edges (or a truncated tree): the atomic climb, by this CTA
        for (size_t s = tid; s < n; s += BTC_THREADS) {
            if (count[s] == 0) continue;
            const unsigned c = gstart[s];
            if (c < m) bh_climb_chain<DIMS>(nodes, c, total, m, arrive);
        }
        return;
    }
    for (unsigned h = tid; h < (unsigned)BTC_HASH; h += BTC_THREADS) sm.hkey[h] = 0u;
    for (unsigned k = tid; k < (E + 1u) * NCHILD; k += BTC_THREADS) sm.kids[k] = 0xffffu;
    if (tid == 0) { sm.maxd = 0; sm.bad = 0; sm.cell[E] = 0u; sm.info[E] = 0u; sm.pid[E] = 0xffffu; }   // node E: the root (cell 0, depth 0)
    __syncthreads();
    // nodes, hash of their cells, data of the complete ones
    for (unsigned e = tid; e < E; e += BTC_THREADS) {
        const uint4 ed = edges[e];
        sm.cell[e] = ed.x; sm.info[e] = ed.z;
        unsigned h = (ed.x * 2654435761u) >> 20;                       // 12 bits
        while (atomicCAS(&sm.hkey[h], 0u, ed.x + 1u) != 0u) h = (h + 1u) & (BTC_HASH - 1u);
        sm.hval[h] = (unsigned short)e;
        if (ed.z & 65536u) {
            const uint4 a = __ldcg(nodes.aux(ed.x));
            const float4 d0 = __ldcg(nodes.data(ed.x));
            sm.val[e] = make_float4(d0.x, d0.y, (DIMS == 3) ? __uint_as_float(a.x) : 0.f, d0.z);
            sm.sub[e] = (a.y ? a.y : total) - ed.x;
        }
        atomicMax(&sm.maxd, (int)((ed.z >> 8) & 0xffu));
    }
    __syncthreads();
    // every edge finds its parent's node and enters itself as the child of its quadrant
    for (unsigned e = tid; e < E; e += BTC_THREADS) {
        const unsigned parent = edges[e].y;
        unsigned pnode = E;                                            // cell 0 is the root
        if (parent != 0u) {
            unsigned h = (parent * 2654435761u) >> 20, probes = 0;
            while (sm.hkey[h] != parent + 1u && sm.hkey[h] != 0u && probes < (unsigned)BTC_HASH) { h = (h + 1u) & (BTC_HASH - 1u); ++probes; }
            if (sm.hkey[h] == parent + 1u) pnode = sm.hval[h];
            else { sm.bad = 1; pnode = 0xffffu; }                      // cannot happen with a consistent list
        }
        sm.pid[e] = (unsigned short)pnode;
        if (pnode != 0xffffu) sm.kids[pnode * NCHILD + (sm.info[e] & 0xffu)] = (unsigned short)e;
    }
    __syncthreads();
    const int top = sm.maxd;
    if (sm.bad) return;
    // level by level: the incomplete cells of depth d - 1 sum their children (all of depth d, all finished)
    for (int d = top; d >= 1; --d) {
        for (unsigned e = tid; e <= E; e += BTC_THREADS) {
            const unsigned inf = sm.info[e];
            if ((int)((inf >> 8) & 0xffu) != d - 1 || (inf & 65536u)) continue;
            float px = 0.f, py = 0.f, pz = 0.f, ms = 0.f;
            unsigned cells = 1;
#pragma unroll
            for (unsigned c = 0; c < NCHILD; ++c) {
                const unsigned kid = sm.kids[e * NCHILD + c];
                if (kid != 0xffffu) {                                  // children in quadrant order
                    const float4 ch = sm.val[kid];
                    px = __fadd_rn(px, __fmul_rn(ch.x, ch.w));
                    py = __fadd_rn(py, __fmul_rn(ch.y, ch.w));
                    if (DIMS == 3) pz = __fadd_rn(pz, __fmul_rn(ch.z, ch.w));
                    ms = __fadd_rn(ms, ch.w);
                    cells += sm.sub[kid];
                }
            }
            if (ms > 0.f) {                                            // 1/m correctly rounded, as the reference's division
                const float inv = __frcp_rn(ms);
                px = __fmul_rn(px, inv);
                py = __fmul_rn(py, inv);
                if (DIMS == 3) pz = __fmul_rn(pz, inv);
            }
            sm.val[e] = make_float4(px, py, pz, ms);
            sm.sub[e] = cells;
            const unsigned c = sm.cell[e];
            if (c < m) {
                __stcg(reinterpret_cast<float2 *>(nodes.data(c)), make_float2(px, py));
                __stcg(reinterpret_cast<float *>(nodes.data(c)) + 2, ms);
                if (DIMS == 3) __stcg(reinterpret_cast<float *>(nodes.aux(c)), pz);
                reinterpret_cast<unsigned *>(nodes.aux(c))[1] = (c + cells < total) ? c + cells : 0u;
            }
        }
        __syncthreads();
    }
    if (TRACE && tid == 0) { trace[0] = E; trace[1] = (unsigned long long)top; trace[3] = (unsigned long long)(clock64() - trace_t0); }
}

// ---- the whole build as ONE cluster kernel (small scenes) ----------------------------------------------------------------
// Same tree, same bits as the launch-per-phase build above (tests compare both with the oracle); what changes is the
// execution model: one thread-block cluster walks through the phases, separated by hardware cluster barriers.
//   box -> root | keys | 8-digit stable radix sort (cluster_prims.cuh) | cells per body + scan -> node offsets |
//   chains: every body emits the cells of its path that first appear with it | per CELL: skip pointer, parent and
//   quadrant (gallop + bisect on the sorted keys -- one thread per cell, so the few large cells near the root no
//   longer serialise in the thread of body 0) | centres of mass level by level, deepest first: a finished cell
//   deposits (position, mass) in its parent's slot for its quadrant, the next level sums its occupied slots in
//   quadrant order -- Quadtree::propagate's arithmetic and order, with a cluster barrier per level instead of
//   atomics-and-fences per cell.
constexpr size_t BH_CL_RANKS_OFFSET = (sizeof(ClSmem) + 15) & ~(size_t)15;
constexpr size_t BH_CL_SMEM = BH_CL_RANKS_OFFSET + (size_t)CL_MAX_CHUNK * sizeof(unsigned short);   // ~115 KB: one CTA per SM
struct BhClusterArgs {
    const float *posm;
    unsigned n, cap;
    BhRoot *root;
    unsigned long long *keys_a, *keys_b;      // the sorted keys end up in keys_a / idx_a
    unsigned *idx_a, *idx_b;
    unsigned *count, *offs;                   // n + 1 words each
    unsigned char *first, *leaf;
    unsigned *owner;                          // per cell: the sorted body that owns it
    unsigned *occ;                            // per cell: occupied quadrants (bit q)
    BhNodes nodes;
    unsigned *status;
    long long *trace;                         // tuning (NBODY_CLUSTER_TRACE): SM clock of CTA 0 after every phase, else null
};

template <int DIMS>
__global__ void __launch_bounds__(CL_THREADS, 1) bh_build_cluster_kernel(const BhClusterArgs a)
{
    int trace_k = 0;
#define BH_TRACE() do { if (a.trace && blockIdx.x == 0 && threadIdx.x == 0) a.trace[trace_k] = clock64(); ++trace_k; } while (0)
    constexpr int BITS = BhT<DIMS>::BITS, LEVELS = BhT<DIMS>::LEVELS;
    constexpr unsigned NCHILD = BhT<DIMS>::NCHILD;
    extern __shared__ __align__(16) unsigned char cl_smem_raw[];
    ClSmem &sm = *reinterpret_cast<ClSmem *>(cl_smem_raw);
    unsigned short *ranks = reinterpret_cast<unsigned short *>(cl_smem_raw + BH_CL_RANKS_OFFSET);
    const unsigned rank = cl_rank(), nc = cl_size(), tid = threadIdx.x, lane = tid & 31;
    const unsigned gtid = rank * CL_THREADS + tid, gthreads = nc * CL_THREADS;
    const unsigned n = a.n;
    BH_TRACE();   // 0: start

    // ---- bounding box -> root quad (Quad::new_containing, Quad.hpp:40-44); same order-preserving keys as bh_bbox_kernel
    {
        unsigned lo[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, hi[3] = {0u, 0u, 0u};
        for (unsigned i = gtid; i < n; i += gthreads) {
            const size_t g = blk_index(i, 0);
#pragma unroll
            for (int c = 0; c < DIMS; ++c) { const unsigned k = f2ord(a.posm[g + c * BLK]); lo[c] = min(lo[c], k); hi[c] = max(hi[c], k); }
        }
        if (tid < 6) sm.xchg[8 + tid] = (tid < 3) ? 0xffffffffu : 0u;
        __syncthreads();
#pragma unroll
        for (int c = 0; c < DIMS; ++c) {
            lo[c] = __reduce_min_sync(0xffffffffu, lo[c]);
            hi[c] = __reduce_max_sync(0xffffffffu, hi[c]);
            if (lane == 0) { atomicMin(&sm.xchg[8 + c], lo[c]); atomicMax(&sm.xchg[11 + c], hi[c]); }
        }
        cl_sync();
        if (tid < 6) {
            unsigned v = (tid < 3) ? 0xffffffffu : 0u;
            for (unsigned c = 0; c < nc; ++c) { const unsigned r = cl_ld_u32(&sm.xchg[8 + tid], c); v = (tid < 3) ? min(v, r) : max(v, r); }
            sm.misc[4 + tid] = (tid < 3) ? ~v : v;          // bh_root_from_box expects the minima complemented
        }
        __syncthreads();
    }
    const BhRoot root = bh_root_from_box<DIMS>(&sm.misc[4]);
    if (gtid == 0) *a.root = root;
    BH_TRACE();   // 1: box

    // ---- quadrant-path keys; clear the occupancy words
    for (unsigned i = gtid; i < n; i += gthreads) {
        const size_t g = blk_index(i, 0);
        const float x = a.posm[g], y = a.posm[g + BLK], z = (DIMS == 3) ? a.posm[g + 2 * BLK] : 0.f;
        float cx = root.cx, cy = root.cy, cz = root.cz, size = root.size;
        unsigned long long k = 0;
#pragma unroll 4
        for (int l = 0; l < LEVELS; ++l) k = (k << BITS) | bh_descend<DIMS>(x, y, z, cx, cy, cz, size);
        k <<= BhT<DIMS>::ALIGN;
        __stcg(a.keys_a + i, k);
        __stcg(a.idx_a + i, i);
    }
    for (unsigned c = gtid; c < a.cap; c += gthreads) a.occ[c] = 0u;
    cl_sync();
    BH_TRACE();   // 1: keys

    // ---- stable sort of (key, body); the result is brought back to the a-buffers
    unsigned long long *keys = a.keys_a;
    unsigned *idx = a.idx_a;
    if (cl_radix_sort<true>(sm, ranks, a.keys_a, a.keys_b, a.idx_a, a.idx_b, n, 0, 64)) {
        for (unsigned i = gtid; i < n; i += gthreads) { __stcg(a.keys_a + i, __ldcg(a.keys_b + i)); __stcg(a.idx_a + i, __ldcg(a.idx_b + i)); }
        cl_sync();
    }

    BH_TRACE();   // 2: sort
    // ---- cells owned by each sorted body (see bh_count_kernel), deepest leaf, exclusive scan -> node offsets
    unsigned maxd = 0;
    for (unsigned s = gtid; s < n; s += gthreads) {
        const unsigned long long k = __ldcg(keys + s);
        unsigned cnt = 0, firstd = 0, leafd = 0;
        if (!(s > 0 && __ldcg(keys + s - 1) == k)) {
            const int lp = (s > 0) ? lcp_levels<DIMS>(__ldcg(keys + s - 1), k) : -1;
            unsigned t = s + 1;
            while (t < n && __ldcg(keys + t) == k) ++t;                  // skip the run of coincident bodies
            const int ln = (t < n) ? lcp_levels<DIMS>(k, __ldcg(keys + t)) : -1;
            leafd = (lp < 0 && ln < 0) ? 0u : (unsigned)(max(lp, ln) + 1);
            firstd = (unsigned)(lp + 1);
            cnt = leafd - firstd + 1;
        }
        a.count[s] = cnt; a.first[s] = (unsigned char)firstd; a.leaf[s] = (unsigned char)leafd;
        maxd = max(maxd, leafd);
    }
    if (gtid == 0) a.count[n] = 0;
    if (tid == 0) sm.xchg[1] = 0;
    __syncthreads();
    maxd = __reduce_max_sync(0xffffffffu, maxd);
    if (lane == 0) atomicMax(&sm.xchg[1], maxd);
    cl_sync();                                                            // count[] complete, per-CTA maxima published
    {
        unsigned m = 0;
        for (unsigned c = 0; c < nc; ++c) m = max(m, cl_ld_u32(&sm.xchg[1], c));
        maxd = m;
    }
    BH_TRACE();   // 3: count
    const unsigned ncells = cl_excl_scan(sm, a.count, a.offs, n + 1);     // offs[n] = number of cells; ends with a barrier
    BH_TRACE();   // 4: scan
    const unsigned m_cells = min(ncells, a.cap);
    if (gtid == 0 && ncells > a.cap && a.status) *reinterpret_cast<volatile unsigned *>(a.status) = 1u;

    // ---- chains: position / mass of the leaf, size^2, cell geometry, owner
    for (unsigned s = gtid; s < n; s += gthreads) {
        const unsigned cnt = a.count[s];
        if (cnt == 0) continue;
        const unsigned long long k = __ldcg(keys + s);
        const unsigned body = __ldcg(idx + s);
        const size_t g = blk_index(body, 0);
        const float x = a.posm[g], y = a.posm[g + BLK], z = (DIMS == 3) ? a.posm[g + 2 * BLK] : 0.f;
        float mass = a.posm[g + 3 * BLK];                              // coincident bodies merge in body-index order (insert() :56-60)
        for (unsigned t = s + 1; t < n && __ldcg(keys + t) == k; ++t) mass = __fadd_rn(mass, a.posm[blk_index(__ldcg(idx + t), 0) + 3 * BLK]);
        const int firstd = a.first[s], leafd = a.leaf[s];
        const unsigned off = __ldcg(a.offs + s);
        float cx = root.cx, cy = root.cy, cz = root.cz, size = root.size;
        for (int d = 0; d <= leafd; ++d) {
            if (d >= firstd) {
                const unsigned c = off + (unsigned)(d - firstd);
                if (c < a.cap) {
                    const bool is_leaf = (d == leafd);
                    __stcg(a.nodes.data(c), make_float4(is_leaf ? x : 0.f, is_leaf ? y : 0.f, is_leaf ? mass : 0.f, __fmul_rn(size, size)));
                    a.nodes.quad[c] = make_float4(cx, cy, size, cz);
                    // aux is completed by the per-cell phase; the leaf's z rides in aux.x
                    if (DIMS == 3) __stcg(reinterpret_cast<float *>(a.nodes.aux(c)), is_leaf ? z : 0.f);
                    a.owner[c] = s;
                }
            }
            if (d < leafd) {
                const unsigned q = (unsigned)(k >> (64 - BITS * (d + 1))) & (NCHILD - 1u);
                const float ns = __fmul_rn(size, 0.5f);
                cx = __fadd_rn(cx, __fmul_rn((q & 1u) ? 0.5f : -0.5f, ns));
                cy = __fadd_rn(cy, __fmul_rn((q & 2u) ? 0.5f : -0.5f, ns));
                if (DIMS == 3) cz = __fadd_rn(cz, __fmul_rn((q & 4u) ? 0.5f : -0.5f, ns));
                size = ns;
            }
        }
    }
    cl_sync();
    BH_TRACE();   // 5: chains

    // ---- per cell: skip pointer (first sorted body after the cell's span), parent, quadrant (see bh_emit_kernel)
    for (unsigned c = gtid; c < m_cells; c += gthreads) {
        const unsigned s = __ldcg(a.owner + c);
        const unsigned long long k = __ldcg(keys + s);
        const int firstd = a.first[s], leafd = a.leaf[s];
        const int d = firstd + (int)(c - __ldcg(a.offs + s));
        unsigned nx = 0;
        if (d > 0) {
            const int sh = 64 - BITS * d;
            const unsigned long long p = k >> sh;
            unsigned lo = s + 1, hi = n, step = 1;                       // first j in (s, n) with (keys[j] >> sh) > p
            while (lo < hi) {
                const unsigned probe = (lo + step - 1 < hi) ? lo + step - 1 : hi - 1;
                if ((__ldcg(keys + probe) >> sh) > p) { hi = probe; break; }
                lo = probe + 1;
                step <<= 1;
            }
            while (lo < hi) {
                const unsigned mid = (lo + hi) >> 1;
                if ((__ldcg(keys + mid) >> sh) > p) hi = mid; else lo = mid + 1;
            }
            nx = (lo < n) ? __ldcg(a.offs + lo) : 0u;
        }
        unsigned par = 0xffffffffu;
        if (d > firstd) par = c - 1;
        else if (d == 1) par = 0u;
        else if (d > 1) {
            const int shp = 64 - BITS * (d - 1);
            const unsigned long long pp = k >> shp;
            unsigned lo = 0, hi = s, step = 1;                           // first j in [0, s] with (keys[j] >> shp) >= pp
            while (lo < hi) {
                const unsigned probe = (hi >= lo + step) ? hi - step : lo;
                if ((__ldcg(keys + probe) >> shp) < pp) { lo = probe + 1; break; }
                hi = probe;
                step <<= 1;
            }
            while (lo < hi) {
                const unsigned mid = (lo + hi) >> 1;
                if ((__ldcg(keys + mid) >> shp) >= pp) hi = mid; else lo = mid + 1;
            }
            par = __ldcg(a.offs + lo) + (unsigned)(d - 1 - (int)a.first[lo]);
        }
        const unsigned qd = (d > 0) ? ((unsigned)(k >> (64 - BITS * d)) & (NCHILD - 1u)) : 0u;
        if (par != 0xffffffffu && par < a.cap) atomicOr(&a.occ[par], 1u << qd);
        uint4 *ax = a.nodes.aux(c);
        reinterpret_cast<unsigned *>(ax)[1] = nx;
        reinterpret_cast<unsigned *>(ax)[2] = (unsigned)d | ((d == leafd) ? 256u : 0u) | (qd << 9);
        reinterpret_cast<unsigned *>(ax)[3] = par;
        if (DIMS != 3) reinterpret_cast<unsigned *>(ax)[0] = 0u;
    }
    cl_sync();
    BH_TRACE();   // 6: cells

    // ---- centres of mass, deepest level first (Quadtree::propagate, :236-258)
    for (int lvl = (int)maxd; lvl >= 0; --lvl) {
        for (unsigned s = gtid; s < n; s += gthreads) {
            if (a.count[s] == 0) continue;
            const int firstd = a.first[s], leafd = a.leaf[s];
            if (lvl < firstd || lvl > leafd) continue;
            const unsigned c = __ldcg(a.offs + s) + (unsigned)(lvl - firstd);
            if (c >= m_cells) continue;
            float x, y, z = 0.f, mass;
            const uint4 ax = __ldcg(a.nodes.aux(c));
            if (lvl == leafd) {
                const float4 d0 = __ldcg(a.nodes.data(c));
                x = d0.x; y = d0.y; mass = d0.z;
                if (DIMS == 3) z = __uint_as_float(ax.x);
            } else {
                const unsigned mask = __ldcg(a.occ + c);
                float4 ch[NCHILD];
#pragma unroll
                for (unsigned q = 0; q < NCHILD; ++q)
                    ch[q] = ((mask >> q) & 1u) ? __ldcg(&a.nodes.slots[(size_t)c * NCHILD + q]) : make_float4(0.f, 0.f, 0.f, 0.f);
                float px = 0.f, py = 0.f, pz = 0.f, ms = 0.f;
#pragma unroll
                for (unsigned q = 0; q < NCHILD; ++q) {
                    if ((mask >> q) & 1u) {                              // children in quadrant order
                        px = __fadd_rn(px, __fmul_rn(ch[q].x, ch[q].w));
                        py = __fadd_rn(py, __fmul_rn(ch[q].y, ch[q].w));
                        if (DIMS == 3) pz = __fadd_rn(pz, __fmul_rn(ch[q].z, ch[q].w));
                        ms = __fadd_rn(ms, ch[q].w);
                    }
                }
                if (ms > 0.f) {                                          // Vec2::operator/=: inv = 1/scalar ; x *= inv ; y *= inv
                    const float inv = __fdiv_rn(1.0f, ms);
                    px = __fmul_rn(px, inv);
                    py = __fmul_rn(py, inv);
                    if (DIMS == 3) pz = __fmul_rn(pz, inv);
                }
                float4 dp = __ldcg(a.nodes.data(c));
                dp.x = px; dp.y = py; dp.z = ms;
                __stcg(a.nodes.data(c), dp);
                if (DIMS == 3) __stcg(reinterpret_cast<float *>(a.nodes.aux(c)), pz);
                x = px; y = py; z = pz; mass = ms;
            }
            const unsigned par = ax.w;
            if (par != 0xffffffffu && par < m_cells) __stcg(&a.nodes.slots[(size_t)par * NCHILD + ((ax.z >> 9) & (NCHILD - 1u))], make_float4(x, y, z, mass));
        }
        cl_sync();
        BH_TRACE();   // 7 + k: level maxd - k
    }
#undef BH_TRACE
}

// one node record = one 256-bit load (sm_100: LDG.E.256); the array is read-only while the walk runs
__device__ __forceinline__ void bh_load_node(const BhNodes &nodes, unsigned i, float4 &nd, uint4 &na)
{
    unsigned r0, r1, r2, r3;
    asm volatile("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(na.x), "=r"(na.y), "=r"(na.z), "=r"(na.w)
                 : "l"(nodes.rec + 2 * (size_t)i));
    nd = make_float4(__uint_as_float(r0), __uint_as_float(r1), __uint_as_float(r2), __uint_as_float(r3));
}

__device__ __forceinline__ float bh_quake(float number)
{
    const float y = __uint_as_float(0x5f3759dfu - (__float_as_uint(number) >> 1));
    return __fmul_rn(y, __fsub_rn(1.5f, __fmul_rn(__fmul_rn(__fmul_rn(number, 0.5f), y), y)));
}

// ---- 7. walk (Quadtree::acc, :113-155) ------------------------------------------------------------------
// one node's contribution to one target; the reference's per-node arithmetic (Quadtree.hpp:119-127) with the z
// terms appended for the octree.  Returns true when the walk should skip the node's subtree (far node or leaf).
// The opening test: displacement to the node's datum, its square, and whether the node is FAR (size^2 < d^2 theta^2,
// Quadtree.hpp:122).  Same expressions as the reference in refcompat mode; fused multiply-adds otherwise.
template <int DIMS, bool REFCOMPAT>
__device__ __forceinline__ bool bh_open_test(const float4 nd, float ndz, float px, float py, float pz, float t_sq,
                                             float &dx, float &dy, float &dz, float &d_sq)
{
    if (!REFCOMPAT) {
        dx = nd.x - px; dy = nd.y - py; dz = (DIMS == 3) ? ndz - pz : 0.f;
        d_sq = fmaf(dx, dx, dy * dy);
        if (DIMS == 3) d_sq = fmaf(dz, dz, d_sq);
        return nd.w < d_sq * t_sq;
    }
    dx = __fsub_rn(nd.x, px); dy = __fsub_rn(nd.y, py); dz = (DIMS == 3) ? __fsub_rn(ndz, pz) : 0.f;
    d_sq = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
    if (DIMS == 3) d_sq = __fadd_rn(d_sq, __fmul_rn(dz, dz));
    return nd.w < __fmul_rn(d_sq, t_sq);
}

// The contribution of an accepted node (far, or a near leaf when those are included): Quadtree.hpp:123-127.
template <int DIMS, bool REFCOMPAT>
__device__ __forceinline__ void bh_accumulate(float mass, float dx, float dy, float dz, float d_sq, bool contributes, float e_sq,
                                              float &ax, float &ay, float &az)
{
    if (!REFCOMPAT) {
        // accurate-rsqrt mode has no bit-for-bit counterpart in the reference: fused multiply-adds, and a select
        // instead of a branch around the interaction (the large-N walk is instruction-issue-bound, profiles/)
        const float inv = rsqrt_approx(d_sq + e_sq);
        const float s3 = (contributes && d_sq > 0.f) ? mass * inv * inv * inv : 0.f;
        ax = fmaf(dx, s3, ax);
        ay = fmaf(dy, s3, ay);
        if (DIMS == 3) az = fmaf(dz, s3, az);
        return;
    }
    if (contributes && d_sq > 0.f) {
        const float inv = bh_quake(__fadd_rn(d_sq, e_sq));
        const float s3 = __fmul_rn(mass, __fmul_rn(__fmul_rn(inv, inv), inv));
        ax = __fadd_rn(ax, __fmul_rn(dx, s3));
        ay = __fadd_rn(ay, __fmul_rn(dy, s3));
        if (DIMS == 3) az = __fadd_rn(az, __fmul_rn(dz, s3));
    }
}

// one node's contribution to one target; returns true when the walk should skip the node's subtree (far node or leaf)
template <int DIMS, bool REFCOMPAT>
__device__ __forceinline__ bool bh_visit(const float4 nd, float ndz, bool is_leaf, float px, float py, float pz, float t_sq,
                                         float e_sq, int fix_near_leaves, float &ax, float &ay, float &az)
{
    float dx, dy, dz, d_sq;
    const bool far = bh_open_test<DIMS, REFCOMPAT>(nd, ndz, px, py, pz, t_sq, dx, dy, dz, d_sq);
    if (!(far || is_leaf)) return false;
    bh_accumulate<DIMS, REFCOMPAT>(nd.z, dx, dy, dz, d_sq, far || fix_near_leaves, e_sq, ax, ay, az);
    return true;
}

// ---- the per-thread walk ----------------------------------------------------------------------------------------------
// One thread per target, targets in Z-order.  A walk is a chain of DEPENDENT node-record loads -- which record comes
// next is known only after the opening test on the current one -- and at the reference's size (25,000 targets = 782
// warps on 148 SMs) nothing hides their latency: the kernel's time is (longest chain in a warp) x (L2 round trip +
// the opening test), ~107 visits per target on the shipped scene.  Tried in round 2 and rejected (profiles/
// r2_walk_window_experiment.txt): a per-thread window of 4 or 8 consecutive records staged in shared memory by
// cp.async, with and without prefetch of the next window -- 3.3x to 5.4x SLOWER (shared-memory bank conflicts on 78 %
// of the wavefronts, 4 to 8 times the bytes per visit), so each visit stays one 256-bit load.
// fused kick-drift epilogue (one GPU, small scenes): the walk thread integrates its own target -- scattered 4-byte
// accesses, which at a few ten thousand bodies cost less than another launch
struct BhFuse {
    float *posm_next, *vel, *acc;
    float G;
    IntegParams ip;
    int insert;          // the collision pass follows: enter the new position into its screening grid (collide.cuh)
    ColArgs ca;
    ColGrid cg;
};

constexpr int WALK_THREADS = 128;

// end of a target's walk: store the acceleration, or -- fused -- integrate the target right away.
// Simulation::iterate after attract(): the same integrate_body_f32 as the stand-alone integrator, same operand order.
template <int DIMS, bool FUSE>
__device__ __forceinline__ void bh_walk_finish(const float *__restrict__ posm, size_t g, unsigned body, size_t shard_start, float px,
                                               float py, float pz, float ax, float ay, float az, float *__restrict__ accp, const BhFuse &fz,
                                               float &new_x, float &new_y)
{
    if (!FUSE) {
        const size_t l = blk_index(body - shard_start, 0);
        accp[l] = ax; accp[l + BLK] = ay; accp[l + 2 * BLK] = az;
        return;
    }
    const float gx = ax * fz.G, gy = ay * fz.G, gz = az * fz.G;
    float qx = px, qy = py, qz = (DIMS == 3) ? pz : posm[g + 2 * BLK];
    float vx = fz.vel[g], vy = fz.vel[g + BLK], vz = fz.vel[g + 2 * BLK];
    integrate_body_f32(qx, qy, qz, vx, vy, vz, gx, gy, gz, fz.ip);
    fz.posm_next[g] = qx; fz.posm_next[g + BLK] = qy; fz.posm_next[g + 2 * BLK] = qz; fz.posm_next[g + 3 * BLK] = posm[g + 3 * BLK];
    new_x = qx; new_y = qy;
    fz.vel[g] = vx; fz.vel[g + BLK] = vy; fz.vel[g + 2 * BLK] = vz;
    fz.acc[g] = gx; fz.acc[g + BLK] = gy; fz.acc[g + 2 * BLK] = gz;
}

template <int DIMS, bool REFCOMPAT, bool FUSE>
__global__ void __launch_bounds__(WALK_THREADS)
bh_walk_direct_kernel(const float *__restrict__ posm, const unsigned *__restrict__ idx, size_t n, BhNodes nodes,
                      float t_sq, float e_sq, int fix_near_leaves, size_t shard_start, size_t shard_count,
                      float *__restrict__ accp, unsigned cap, unsigned long long *visits, const BhFuse fz,
                      const unsigned *__restrict__ n_dev)
{
    pdl_enter();
    if (n_dev) n = min(n, (size_t)*n_dev);                  // sharded: `idx` is the compacted list of this GPU's targets
    const size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned body = s < n ? idx[s] : 0u;              // targets in Z-order: neighbouring threads walk alike
    const bool active = s < n && body >= shard_start && body < shard_start + shard_count;
    if (!(FUSE && fz.insert) && !active) return;            // (with the grid insert the whole warp stays: see the end)
    const size_t g = blk_index(body, 0);
    const float px = posm[g], py = posm[g + BLK], pz = (DIMS == 3) ? posm[g + 2 * BLK] : 0.f;
    float ax = 0.f, ay = 0.f, az = 0.f;
    // With one or two warps per scheduler (the reference's 25,000 targets) nothing hides a load: a warp's time is its
    // longest walk x (load round trip + the arithmetic of a visit).  The loop is software-pipelined: the record of the
    // NEXT node is requested as soon as the opening test has chosen it, and the force arithmetic of the current node
    // (the Quake rsqrt chain) runs while that load is in flight.  Same visits, same operations, same order of additions
    // as Quadtree::acc.  (Measured and dropped, profiles/r2_walk_experiments.txt: prefetching the following records
    // into L1, requesting BOTH possible successors before the test, and thin warps -- 16 / 8 / 4 targets per warp so that
    // every scheduler has several warps: none shortens the chain, whose length is the LONGEST walk of the scene, 173
    // visits of ~450 cycles on the shipped scene, tools/bh_walk_stats.py.)
    unsigned i = 0, nvis = 0;
    float4 nd;
    uint4 na;
    bh_load_node(nodes, 0u, nd, na);
    while (active) {
        ++nvis;
        float dx, dy, dz, d_sq;
        const bool far = bh_open_test<DIMS, REFCOMPAT>(nd, __uint_as_float(na.x), px, py, pz, t_sq, dx, dy, dz, d_sq);
        const bool skip = far || (na.z & 256u) != 0u;                    // far node or leaf: its subtree is not entered
        const unsigned nxt = skip ? na.y : i + 1u;
        const bool more = nxt != 0u && nxt < cap;                        // nxt >= cap only if the tree overflowed (status word)
        const float mass = nd.z;
        float4 nd2 = nd;
        uint4 na2 = na;
        if (more) bh_load_node(nodes, nxt, nd2, na2);
        if (skip) bh_accumulate<DIMS, REFCOMPAT>(mass, dx, dy, dz, d_sq, far || fix_near_leaves, e_sq, ax, ay, az);
        if (!more) break;
        nd = nd2; na = na2; i = nxt;
    }
    float new_x = 0.f, new_y = 0.f;
    if (active) {
        if (visits) {                                       // profiled steps only: one atomic pair per warp, not per thread
            const unsigned am = __activemask();
            const unsigned sum = __reduce_add_sync(am, nvis), mx = __reduce_max_sync(am, nvis);
            if ((threadIdx.x & 31u) == (unsigned)(__ffs(am) - 1)) { atomicAdd(visits, (unsigned long long)sum); atomicMax(visits + 1, (unsigned long long)mx); }
        }
        bh_walk_finish<DIMS, FUSE>(posm, g, body, shard_start, px, py, pz, ax, ay, az, accp, fz, new_x, new_y);
    }
    if (FUSE && fz.insert) {
        // Simulation::step(): iterate() ; collide().  The pass starts by entering every body into its screening grid under
        // the cells and strips it covers; the thread that just computed a body's new position does that here instead of a
        // kernel of its own (a body covering many units is shared out over the warp's lanes, hence all 32 stay to the end).
        ColBody b;
        b.x = new_x; b.y = new_y; b.r = active ? fz.vel[g + 3 * BLK] : 0.f;
        unsigned unused = 0;
        col_grid_lane_units(fz.ca, fz.cg, active, b, body, true, ColInsertOp(), unused);
    }
}

// Warp-cooperative walk.  The 32 lanes of a warp hold 32 targets that are neighbours in Z-order; the
// warp walks the UNION of their traversals in pre-order index order and every visited node record is
// loaded once per warp (uniform address -> one broadcast transaction instead of up to 32 divergent
// ones).  Each lane keeps `resume`, the index of the next node of ITS OWN sequential walk (node+1 when
// it opens a branch, next[node] when it accepts a far node or passes a leaf); it takes part only at
// nodes == resume, and the warp advances to the minimum resume over lanes.  A lane's sequence of
// visited nodes, and therefore its sum term by term, is exactly that of Quadtree::acc -- bit-exact
// like bh_walk_kernel, only the memory access pattern differs.
template <int DIMS, bool REFCOMPAT>
__global__ void __launch_bounds__(128)
bh_walk_warp_kernel(const float *__restrict__ posm, const unsigned *__restrict__ idx, size_t n, BhNodes nodes,
                    float t_sq, float e_sq, int fix_near_leaves, size_t shard_start, size_t shard_count,
                    float *__restrict__ accp, unsigned cap, unsigned window, unsigned long long *visits,
                    const unsigned *__restrict__ n_dev)
{
    if (n_dev) n = min(n, (size_t)*n_dev);                  // sharded: `idx` is the compacted list of this GPU's targets
    const size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = s < n;
    const unsigned body = valid ? idx[s] : 0u;
    const bool mine = valid && body >= shard_start && body < shard_start + shard_count;
    const size_t g = blk_index(body, 0);
    const float px = posm[g], py = posm[g + BLK], pz = (DIMS == 3) ? posm[g + 2 * BLK] : 0.f;
    float ax = 0.f, ay = 0.f, az = 0.f;
    constexpr unsigned DONE = 0xffffffffu;
    // (Staging batches of 32 consecutive records in shared memory was tried and is SLOWER, 7.3 vs 6.2 ms at 1M
    //  bodies: the warp-uniform loads below already hit L1, the walk is issue-bound, and a batch is rarely used up
    //  before the walk skips past it.)
    // `window`: a lane takes its next node whenever that node lies fewer than `window` records ahead of the
    // slowest lane, so one step's loads fall into `window` consecutive 32-byte records (window = 1: every active
    // lane reads the same record).  Lanes that accepted a cell keep going while a neighbour descends into it.
    unsigned resume = mine ? 0u : DONE, nvis = 0;
    unsigned i = __reduce_min_sync(0xffffffffu, resume);
    while (i < cap) {                                        // DONE (and an overflowed tree) end the walk
        if (resume - i < window) {                           // unsigned: false for DONE
            float4 nd;
            uint4 na;
            bh_load_node(nodes, resume, nd, na);
            ++nvis;
            if (bh_visit<DIMS, REFCOMPAT>(nd, __uint_as_float(na.x), (na.z & 256u) != 0u, px, py, pz, t_sq, e_sq, fix_near_leaves, ax, ay, az))
                resume = na.y ? na.y : DONE;
            else
                resume = resume + 1;
        }
        i = __reduce_min_sync(0xffffffffu, resume);
    }
    if (mine) {
        const size_t l = blk_index(body - shard_start, 0);
        accp[l] = ax; accp[l + BLK] = ay; accp[l + 2 * BLK] = az;
    }
    if (visits && nvis) { atomicAdd(visits, (unsigned long long)nvis); atomicMax(visits + 1, (unsigned long long)nvis); }
}

// Several GPUs: every GPU builds the whole tree but walks only the targets of its shard (a range of BODY indices,
// scattered all over the Z-order).  Compact them IN Z-ORDER -- a chained scan over the blocks of 256 sorted bodies (ticket,
// look-back over the predecessors' status words, cf. scan_excl_kernel) -- so that the walk's warps are full of this GPU's
// targets and every warp's 32 targets are neighbours in space.  (The first version kept the order inside a block and took
// the blocks in arrival order: with 8 GPUs a block holds ~32 of a GPU's targets, so nearly every warp straddled two
// blocks from unrelated places and the warp-cooperative walk paid for the union of both: 3.4 ms instead of 1.9 ms
// for an eighth of 4M targets.)  `scratch`: ticket | status words, zeroed by the build's memset; the last block
// leaves the number of targets in `counter`.
__global__ void __launch_bounds__(256)
bh_shard_targets_kernel(const unsigned *__restrict__ idx, size_t n, size_t shard_start, size_t shard_count,
                        unsigned *__restrict__ out, unsigned *__restrict__ counter, unsigned *__restrict__ ticket, unsigned long long *status)
{
    __shared__ unsigned warp_sums[RS_WARPS];
    __shared__ unsigned s_tile, s_base;
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const unsigned tile = s_tile;
    const size_t s = (size_t)tile * 256 + tid;
    const unsigned body = s < n ? idx[s] : 0u;
    const unsigned mine = (s < n && body >= shard_start && body < shard_start + shard_count) ? 1u : 0u;
    unsigned total = 0;
    const unsigned before = rs_block_excl_scan(mine, warp_sums, &total);
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
            const int firstp = mask ? __ffs(mask) - 1 : 31;
            unsigned c = (lane <= firstp) ? (unsigned)v : 0u;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
            excl += c;
            if (mask) break;
            t -= 32;
        }
        if (lane == 0) {
            s_base = excl;
            if (tile > 0) rs_st_status(status + tile, rs_pack(2u, excl + total));
            if ((size_t)(tile + 1) * 256 >= n) *counter = excl + total;      // the last block: this GPU's number of targets
        }
    }
    __syncthreads();
    if (mine) out[s_base + before] = body;
}

// ---- host side ------------------------------------------------------------------------------------------------
// Cells: a body owns the cells of its path that first appear with it -- about 2.8 n for well-separated bodies (the
// reference's shipped scene: 69,681 cells for 25,000 bodies), up to 32 n when many bodies are much closer to a
// neighbour than to everything else (each such pair hangs from a chain of single-child cells).  `node_factor` x n
// cells are reserved (default 4; environment NBODY_BH_NODE_FACTOR); a tree that needs more raises the context's
// sticky status instead of being truncated silently.
// cluster_mode: 0 = the single-cluster build for scenes of up to cluster_max_n bodies (if the device can host the
// cluster), 1 = never, 2 = required (cudaErrorNotSupported otherwise)
cudaError_t BhWorkspace::alloc(size_t n, int dims_, double node_factor, int cluster_mode, size_t cluster_max_n)
{
    cudaError_t e;
    n_cap = n;
    dims = dims_;
    cluster_ctas = 0;
    if (cluster_mode != 1) {
        const int avail = cluster_ctas_available(dims_);
        const size_t lim = std::min<size_t>((size_t)avail * CL_MAX_CHUNK, cluster_mode == 2 ? (size_t)-1 : cluster_max_n);
        if (avail > 0 && n <= lim) cluster_ctas = avail;
        else if (cluster_mode == 2) return cudaErrorNotSupported;
    }
    node_cap = (unsigned)std::min<size_t>((size_t)(std::max(1.0, std::min(node_factor, 33.0)) * (double)n) + 1024, 0x7fffffffu);
#define BH_ALLOC(p, bytes) if ((e = cudaMalloc((void **)&(p), (bytes))) != cudaSuccess) return e;
    BH_ALLOC(root, sizeof(BhRoot))
    BH_ALLOC(keys_in, n * 8) BH_ALLOC(keys, n * 8) BH_ALLOC(idx_in, n * 4) BH_ALLOC(idx, n * 4)
    BH_ALLOC(count, (n + 2) * 4) BH_ALLOC(offs, (n + 2) * 4) BH_ALLOC(first, n) BH_ALLOC(leaf, n)
    BH_ALLOC(node_data, ((size_t)node_cap + 16) * 32) BH_ALLOC(node_quad, (size_t)node_cap * 16)
    BH_ALLOC(node_slots, (size_t)node_cap * (dims == 3 ? 8 : 4) * 16)
    BH_ALLOC(node_slot_cells, (size_t)node_cap * (dims == 3 ? 8 : 4) * 4)
    BH_ALLOC(node_owner, (size_t)node_cap * 4)
    BH_ALLOC(shard_targets, (n + 1) * 4)
    BH_ALLOC(climb_start, (n + 1) * 4)
    // everything that must be zero at the start of a build lives in ONE region cleared by one memset per step:
    // bounding box | radix-sort scratch (histograms, tickets, status words) | scan scratch | arrival counters
    auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
    // ... | scan scratch of the target compaction of a sharded walk
    const size_t o_sort = 256, o_scan = o_sort + up(radix_sort_temp_bytes(n)), o_shard = o_scan + up(exclusive_scan_temp_bytes(n + 1));
    const size_t o_arrive = o_shard + up(64 + ((n + 255) / 256 + 1) * sizeof(unsigned long long));
    zero_bytes = o_arrive + (size_t)node_cap * 4;
    BH_ALLOC(zero_region, zero_bytes)
    box = zero_region;
    sort_temp = (char *)zero_region + o_sort;
    scan_temp = (char *)zero_region + o_scan;
    shard_scan = (char *)zero_region + o_shard;
    node_arrive = (char *)zero_region + o_arrive;
#undef BH_ALLOC
    return cudaSuccess;
}

void BhWorkspace::release()
{
    void *ptrs[] = {root, keys_in, keys, idx_in, idx, count, offs, first, leaf, node_data, node_quad, node_slots, node_slot_cells, node_owner, shard_targets, climb_start, zero_region, trace};
    for (void *p : ptrs) if (p) cudaFree(p);
    *this = BhWorkspace();
}

static BhNodes bh_nodes(const BhWorkspace &w)
{
    BhNodes nd;
    nd.rec = (float4 *)w.node_data; nd.quad = (float4 *)w.node_quad; nd.slots = (float4 *)w.node_slots; nd.slot_cells = (unsigned *)w.node_slot_cells;
    return nd;
}

template <int DIMS>
static cudaError_t bh_build_t(BhWorkspace &w, const float *posm, size_t n, cudaStream_t st, int *launches)
{
    cudaError_t e;
    const unsigned g256 = (unsigned)((n + 255) / 256), g128 = (unsigned)((n + 127) / 128);
    unsigned *cnt = (unsigned *)w.count;
    // one memset: bounding box, sort and scan scratch, arrival counters (at most min(node_cap, 4n + 1024) cells)
    if ((e = cudaMemsetAsync(w.zero_region, 0, w.zero_bytes, st)) != cudaSuccess) return e;
    bh_bbox_kernel<DIMS><<<std::min(g256, 4u * 148u), 256, 0, st>>>(posm, n, (unsigned *)w.box);
    // stable LSD sort; the result lands back in the first buffer pair -> swap roles
    // (from here on a kernel may become resident while its predecessor still runs: launch_pdl / pdl_enter, common.cuh)
    if ((e = launch_pdl(bh_keys_kernel<DIMS, true>, dim3(g256), dim3(256), 0, st, posm, n, (const unsigned *)w.box, (BhRoot *)w.root,
                        (unsigned long long *)w.keys_in, (unsigned *)w.idx_in, (unsigned *)w.sort_temp)) != cudaSuccess) return e;
    // high digits first: the leading 32 bits (16 quadtree levels; 48 bits = 16 octree levels) order all bodies but the few
    // that share a deeper cell, which the sort's run repair puts right (radix_sort.cuh)
    if ((e = radix_sort_u64((unsigned long long *)w.keys_in, (unsigned long long *)w.keys, (unsigned *)w.idx_in, (unsigned *)w.idx,
                            n, w.sort_temp, st, 0, 64, launches, nullptr, true, DIMS == 3 ? 16 : 32)) != cudaSuccess) return e;
    std::swap(w.keys_in, w.keys);
    std::swap(w.idx_in, w.idx);
    if ((e = launch_pdl(bh_count_kernel<DIMS>, dim3(g256), dim3(256), 0, st, (const unsigned long long *)w.keys, n, cnt, (unsigned char *)w.first,
                        (unsigned char *)w.leaf)) != cudaSuccess) return e;
    if ((e = exclusive_scan_u32((const unsigned *)w.count, (unsigned *)w.offs, n + 1, w.scan_temp, st, launches, true)) != cudaSuccess) return e;
    // cells, centres of mass and skip pointers: the local kernel + the climb over the few long cells (default), or the
    // round-1/2 pair emit + climb-from-the-leaves (NBODY_BH_LOCAL=0; bit-identical trees, kept for comparison)
    static const bool local_off = getenv("NBODY_BH_LOCAL") && atoi(getenv("NBODY_BH_LOCAL")) == 0;
    if (!local_off) {
        static const bool want_trace = getenv("NBODY_BH_TRACE") != nullptr;
        static const int geom = getenv("NBODY_BHL_GEOM") ? atoi(getenv("NBODY_BHL_GEOM")) : 0;   // tuning: bodies per CTA / halo
        // small scenes: the (child, parent) edges left to the climb are listed (word 8 of the zeroed box region counts them) and
        // ONE CTA finishes the top of the tree in shared memory; NBODY_BH_CTA_CLIMB=0 keeps the atomic climb at every size
        static const bool cta_climb_off = getenv("NBODY_BH_CTA_CLIMB") && atoi(getenv("NBODY_BH_CTA_CLIMB")) == 0;
        if (w.top_cta_state == 0) {                          // once per workspace, i.e. per device: function attributes are per device
            const bool ok = cudaFuncSetAttribute(bh_top_cta_kernel<DIMS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BtcSmem<DIMS>)) == cudaSuccess &&
                            cudaFuncSetAttribute(bh_top_cta_kernel<DIMS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BtcSmem<DIMS>)) == cudaSuccess;
            if (!ok) cudaGetLastError();
            w.top_cta_state = ok ? 1 : -1;
        }
        const bool cta_climb = !cta_climb_off && w.top_cta_state > 0 && n <= 32768;
        // (tests force the kernel's fall-back with a small limit)
        static const unsigned top_edge_limit = getenv("NBODY_BH_TOP_MAX_EDGES") ? (unsigned)std::min(atoi(getenv("NBODY_BH_TOP_MAX_EDGES")), BTC_MAX_EDGES) : (unsigned)BTC_MAX_EDGES;
#define BHL_ARGS posm, (const unsigned long long *)w.keys, (const unsigned *)w.idx, n, (const BhRoot *)w.root, (const unsigned *)w.offs,      \
                 (const unsigned *)w.count, (const unsigned char *)w.first, (const unsigned char *)w.leaf, bh_nodes(w), (unsigned *)w.node_arrive, \
                 w.node_cap, w.status, (unsigned *)w.climb_start, cta_climb ? (uint4 *)w.node_owner : nullptr, (unsigned *)w.box + 8, w.node_cap / 4
#define BHL_LAUNCH(T, M, H, TR) do { if ((e = launch_pdl(bh_emit_local_kernel<DIMS, T, M, H>, dim3((unsigned)((n + (M) - 1) / (M))), dim3((M) + (H)), 0, st, \
                                                                 BHL_ARGS, TR)) != cudaSuccess) return e; } while (0)
        if (want_trace) {
            const unsigned lgrid_trace = (unsigned)((n + 255) / 256);
            if (!w.trace && cudaMalloc(&w.trace, 2048 * sizeof(long long)) != cudaSuccess) return cudaErrorMemoryAllocation;
            cudaMemsetAsync(w.trace, 0, 2048 * sizeof(long long), st);
            BHL_LAUNCH(true, 256, 128, (unsigned long long *)w.trace);
            unsigned long long h[2048];
            if (cudaMemcpyAsync(h, w.trace, sizeof h, cudaMemcpyDeviceToHost, st) == cudaSuccess && cudaStreamSynchronize(st) == cudaSuccess) {
                if (getenv("NBODY_BH_TRACE_CTAS")) {
                    fprintf(stderr, "[emit_local] per CTA (cycles in the level loop / iterations):");
                    for (unsigned c = 0; c < lgrid_trace && c < 1000; ++c) fprintf(stderr, " %llu/%llu", h[16 + 2 * c], h[17 + 2 * c]);
                    fprintf(stderr, "\n");
                }
                fprintf(stderr, "[emit_local n=%zu] max cycles to the end of: loads %llu, skeleton %llu, local levels %llu, hand-over %llu; level iterations: CTA 0 %llu, mean of the others %.1f, most %llu\n",
                        n, h[0], h[1], h[2], h[3], h[8], (double)h[9] / (double)std::max<size_t>(1, (n + 255) / 256 - 1), h[10]);
            }
        } else if (geom == 1) BHL_LAUNCH(false, 128, 128, nullptr);
        else if (geom == 2) BHL_LAUNCH(false, 64, 64, nullptr);
        else if (geom == 3) BHL_LAUNCH(false, 128, 256, nullptr);
        else if (geom == 5) BHL_LAUNCH(false, 64, 192, nullptr);
        else if (geom == 6) BHL_LAUNCH(false, 512, 512, nullptr);
        else if (geom == 7) BHL_LAUNCH(false, 256, 256, nullptr);
        else BHL_LAUNCH(false, 256, 128, nullptr);           // sweep on B200 (profiles/): best at 1M and 4M bodies, no difference at 25,000
#undef BHL_LAUNCH
#undef BHL_ARGS
#define BCC_ARGS bh_nodes(w), n, (const unsigned *)w.offs, (const unsigned *)w.count, (const uint4 *)w.node_owner, (const unsigned *)w.box + 8, w.node_cap / 4, \
                 (const unsigned *)w.climb_start, (unsigned *)w.node_arrive, w.node_cap, top_edge_limit
        if (want_trace) {
            unsigned long long *tr = (unsigned long long *)w.trace + 32;
            cudaMemsetAsync(tr, 0, 4 * sizeof(long long), st);
            if (cta_climb) e = launch_pdl(bh_top_cta_kernel<DIMS, true>, dim3(1), dim3(BTC_THREADS), sizeof(BtcSmem<DIMS>), st, BCC_ARGS, tr);
            else bh_climb_kernel<DIMS, true><<<g256, 256, 0, st>>>(bh_nodes(w), n, (const unsigned *)w.offs, (const unsigned *)w.count,
                                                                   (const unsigned *)w.climb_start, (unsigned *)w.node_arrive, w.node_cap, tr);
            unsigned long long h[4];
            if (cudaMemcpyAsync(h, tr, sizeof h, cudaMemcpyDeviceToHost, st) == cudaSuccess && cudaStreamSynchronize(st) == cudaSuccess)
                fprintf(stderr, cta_climb ? "[top of the tree, one CTA, n=%zu] edges %llu, deepest %llu, (%llu), cycles %llu\n"
                                          : "[climb n=%zu] chains %llu, most cells finished by one thread %llu (%llu with several children), max cycles %llu\n", n, h[0], h[1], h[2], h[3]);
        } else if (cta_climb)
            e = launch_pdl(bh_top_cta_kernel<DIMS, false>, dim3(1), dim3(BTC_THREADS), sizeof(BtcSmem<DIMS>), st, BCC_ARGS, (unsigned long long *)nullptr);
        else
            e = launch_pdl(bh_climb_kernel<DIMS, false>, dim3(g256), dim3(256), 0, st, bh_nodes(w), n, (const unsigned *)w.offs, (const unsigned *)w.count,
                           (const unsigned *)w.climb_start, (unsigned *)w.node_arrive, w.node_cap, (unsigned long long *)nullptr);
#undef BCC_ARGS
        if (e != cudaSuccess) return e;
    } else {
        bh_emit_kernel<DIMS><<<g128, 128, 0, st>>>(posm, (const unsigned long long *)w.keys, (const unsigned *)w.idx, n, (const BhRoot *)w.root,
                                                   (const unsigned *)w.offs, (const unsigned *)w.count, (const unsigned char *)w.first,
                                                   (const unsigned char *)w.leaf, bh_nodes(w), (unsigned *)w.node_arrive, w.node_cap, w.status);
        bh_propagate_kernel<DIMS><<<g256, 256, 0, st>>>(bh_nodes(w), n, (const unsigned *)w.offs, (const unsigned char *)w.first,
                                                        (const unsigned char *)w.leaf, (const unsigned *)w.count, (unsigned *)w.node_arrive,
                                                        w.node_cap);
    }
    w.count_valid = false;
    if (launches) *launches += 5;                            // the sort and the scan count their own launches
    return cudaGetLastError();
}

// How many CTAs the single-cluster kernels run with on this device: 16 (non-portable size) if a cluster of 16 can be
// resident, else 8; 0 if not even that (then only the launch-per-phase path exists).
template <typename K>
static int cluster_ctas_for(K kernel, size_t smem)
{
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return 0; }
    const bool np_ok = cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess;
    if (!np_ok) cudaGetLastError();
    for (int nc = np_ok ? 16 : 8; nc >= 8; nc -= 8) {
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof cfg);
        cfg.gridDim = dim3((unsigned)nc); cfg.blockDim = dim3(CL_THREADS); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at;
        at.id = cudaLaunchAttributeClusterDimension;
        at.val.clusterDim.x = (unsigned)nc; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
        cfg.attrs = &at; cfg.numAttrs = 1;
        int nclusters = 0;
        if (cudaOccupancyMaxActiveClusters(&nclusters, kernel, &cfg) == cudaSuccess && nclusters >= 1) return nc;
        cudaGetLastError();
    }
    return 0;
}

template <typename K, typename A>
static cudaError_t launch_cluster(K kernel, int nc, size_t smem, const A &args, cudaStream_t st)
{
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3((unsigned)nc); cfg.blockDim = dim3(CL_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at;
    at.id = cudaLaunchAttributeClusterDimension;
    at.val.clusterDim.x = (unsigned)nc; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
    cfg.attrs = &at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, args);
}

int BhWorkspace::cluster_ctas_available(int dims_)
{
    return dims_ == 3 ? cluster_ctas_for(bh_build_cluster_kernel<3>, BH_CL_SMEM) : cluster_ctas_for(bh_build_cluster_kernel<2>, BH_CL_SMEM);
}

// Build the tree of the first n bodies of `posm`.  Fully asynchronous (no host read-back): the node count
// stays on the device (offs[n]); n_nodes is fetched lazily by node_count().
cudaError_t BhWorkspace::build(const float *posm, size_t n, cudaStream_t st, int *launches)
{
    if (n == 0 || n > n_cap) return cudaErrorInvalidValue;
    if (cluster_ctas > 0 && n <= (size_t)cluster_ctas * CL_MAX_CHUNK) {        // small scene: the whole build is one cluster kernel
        BhClusterArgs a;
        a.posm = posm; a.n = (unsigned)n; a.cap = node_cap; a.root = (BhRoot *)root;
        a.keys_a = (unsigned long long *)keys_in; a.keys_b = (unsigned long long *)keys;
        a.idx_a = (unsigned *)idx_in; a.idx_b = (unsigned *)idx;
        a.count = (unsigned *)count; a.offs = (unsigned *)offs; a.first = (unsigned char *)first; a.leaf = (unsigned char *)leaf;
        a.owner = (unsigned *)node_owner; a.occ = (unsigned *)node_arrive; a.nodes = bh_nodes(*this); a.status = status;
        static const bool want_trace = getenv("NBODY_CLUSTER_TRACE") != nullptr;
        a.trace = nullptr;
        if (want_trace) {
            if (!trace && cudaMalloc(&trace, 2048 * sizeof(long long)) != cudaSuccess) return cudaErrorMemoryAllocation;
            cudaMemsetAsync(trace, 0, 64 * sizeof(long long), st);
            a.trace = (long long *)trace;
        }
        const cudaError_t e = dims == 3 ? launch_cluster(bh_build_cluster_kernel<3>, cluster_ctas, BH_CL_SMEM, a, st)
                                        : launch_cluster(bh_build_cluster_kernel<2>, cluster_ctas, BH_CL_SMEM, a, st);
        if (e != cudaSuccess) return e;
        if (want_trace) {                                                      // tuning aid: cycles per phase of this build, on stderr
            long long h[64];
            if (cudaMemcpyAsync(h, trace, sizeof h, cudaMemcpyDeviceToHost, st) == cudaSuccess && cudaStreamSynchronize(st) == cudaSuccess) {
                fprintf(stderr, "[cluster build n=%zu ctas=%d] cycles: box..", n, cluster_ctas);
                for (int k = 1; k < 64 && h[k]; ++k) fprintf(stderr, " %lld", h[k] - h[k - 1]);
                fprintf(stderr, "\n");
            }
        }
        std::swap(keys_in, keys);                                              // the sorted pairs are in the a-buffers
        std::swap(idx_in, idx);
        count_valid = false;
        if (launches) *launches += 1;
        return cudaGetLastError();
    }
    return dims == 3 ? bh_build_t<3>(*this, posm, n, st, launches) : bh_build_t<2>(*this, posm, n, st, launches);
}

// number of cells of the last tree (synchronises the stream once, then cached)
cudaError_t BhWorkspace::node_count(size_t n, cudaStream_t st, unsigned *out)
{
    if (!count_valid) {
        unsigned m = 0;
        cudaError_t e;
        if ((e = cudaMemcpyAsync(&m, (unsigned *)offs + n, 4, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
        if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;
        n_nodes = m;
        count_valid = true;
    }
    *out = n_nodes;
    return n_nodes > node_cap ? cudaErrorMemoryAllocation : cudaSuccess;   // pathological depth: more cells than reserved
}

template <int DIMS, bool FUSE>
static void bh_walk_t(const BhWorkspace &w, const float *posm, size_t n, float t_sq, float e_sq, bool refcompat, int fix,
                      size_t shard_start, size_t shard_count, float *accp, unsigned long long *visits, const BhFuse &fz, cudaStream_t st)
{
    const unsigned *idx = (const unsigned *)w.idx;
    const unsigned *n_dev = nullptr;
    const BhNodes nd = bh_nodes(w);
    if (!(shard_start == 0 && shard_count >= n)) {           // this GPU owns a shard: walk the compacted list of its targets
        unsigned *counter = (unsigned *)w.shard_targets + w.n_cap;
        bh_shard_targets_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(idx, n, shard_start, shard_count, (unsigned *)w.shard_targets, counter,
                                                                               (unsigned *)w.shard_scan, (unsigned long long *)((char *)w.shard_scan + 64));
        idx = (const unsigned *)w.shard_targets;
        n_dev = counter;
        n = std::min(n, shard_count);
    }
    const unsigned g = (unsigned)((n + 127) / 128);
    if (w.warp_walk && !FUSE) {
        if (refcompat) bh_walk_warp_kernel<DIMS, true><<<g, 128, 0, st>>>(posm, idx, n, nd, t_sq, e_sq, fix, shard_start, shard_count, accp, w.node_cap, w.walk_window, visits, n_dev);
        else bh_walk_warp_kernel<DIMS, false><<<g, 128, 0, st>>>(posm, idx, n, nd, t_sq, e_sq, fix, shard_start, shard_count, accp, w.node_cap, w.walk_window, visits, n_dev);
        return;
    }
    if (refcompat) launch_pdl(bh_walk_direct_kernel<DIMS, true, FUSE>, dim3(g), dim3(WALK_THREADS), 0, st, posm, idx, n, nd, t_sq, e_sq, fix, shard_start, shard_count, accp, w.node_cap, visits, fz, n_dev);
    else launch_pdl(bh_walk_direct_kernel<DIMS, false, FUSE>, dim3(g), dim3(WALK_THREADS), 0, st, posm, idx, n, nd, t_sq, e_sq, fix, shard_start, shard_count, accp, w.node_cap, visits, fz, n_dev);
}

// `fuse` != nullptr: the walk threads also integrate their targets (kick-drift; one GPU, whole array = one shard)
cudaError_t BhWorkspace::walk(const float *posm, size_t n, float theta, float eps, bool refcompat, bool fix_near_leaves,
                              size_t shard_start, size_t shard_count, float *accp, unsigned long long *visits,
                              const BhFuseArgs *fuse, cudaStream_t st)
{
    const float t_sq = theta * theta, e_sq = eps * eps;       // Quadtree ctor, Quadtree.hpp:19
    const int fix = fix_near_leaves ? 1 : 0;
    BhFuse fz;
    memset(&fz, 0, sizeof fz);
    if (fuse) {
        fz.posm_next = fuse->posm_next; fz.vel = fuse->vel; fz.acc = fuse->acc; fz.G = fuse->G; fz.ip = fuse->ip;
        if (fuse->col_args && fuse->col_grid) { fz.insert = 1; fz.ca = *fuse->col_args; fz.cg = *fuse->col_grid; }
    }
    if (fuse) {
        if (dims == 3) bh_walk_t<3, true>(*this, posm, n, t_sq, e_sq, refcompat, fix, shard_start, shard_count, accp, visits, fz, st);
        else bh_walk_t<2, true>(*this, posm, n, t_sq, e_sq, refcompat, fix, shard_start, shard_count, accp, visits, fz, st);
    } else {
        if (dims == 3) bh_walk_t<3, false>(*this, posm, n, t_sq, e_sq, refcompat, fix, shard_start, shard_count, accp, visits, fz, st);
        else bh_walk_t<2, false>(*this, posm, n, t_sq, e_sq, refcompat, fix, shard_start, shard_count, accp, visits, fz, st);
    }
    return cudaGetLastError();
}

// node array in walk order for the parity tests: f8 = (x, y, z, mass, cx, cy, cz, size), u2 = (next, depth | leaf<<8)
cudaError_t BhWorkspace::download_nodes(float *f8, unsigned *u2, size_t cap, cudaStream_t st)
{
    const size_t m = std::min<size_t>(std::min<size_t>(cap, n_nodes), node_cap);
    std::vector<float4> rec(2 * m), q(m);
    cudaError_t e;
    if ((e = cudaMemcpyAsync(rec.data(), node_data, m * 32, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
    if ((e = cudaMemcpyAsync(q.data(), node_quad, m * 16, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;
    for (size_t i = 0; i < m; ++i) {
        const float4 d = rec[2 * i];
        uint4 a;
        memcpy(&a, &rec[2 * i + 1], 16);
        float z;
        memcpy(&z, &a.x, 4);
        f8[8 * i + 0] = d.x; f8[8 * i + 1] = d.y; f8[8 * i + 2] = (dims == 3) ? z : 0.f; f8[8 * i + 3] = d.z;
        f8[8 * i + 4] = q[i].x; f8[8 * i + 5] = q[i].y; f8[8 * i + 6] = (dims == 3) ? q[i].w : 0.f; f8[8 * i + 7] = q[i].z;
        u2[2 * i + 0] = a.y; u2[2 * i + 1] = a.z & 0x1ffu;   // depth | leaf << 8 (the quadrant bits stay internal)
    }
    return cudaSuccess;
}

} // namespace nb
