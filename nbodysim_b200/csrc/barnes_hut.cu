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
//   3. stable radix sort of (key, body)                                (cub::DeviceRadixSort)
//   4. cells owned by each sorted body = the path cells that first appear with it; exclusive scan
//      -> depth-first pre-order node array WITHOUT the reference's empty leaves (they contribute +-0)
//   5. skip pointers (`next`) by binary search on the sorted keys; first child = index + 1
//   6. centres of mass bottom-up, one launch per level, children summed in quadrant order
//   7. walk: one thread per target (targets in Z-order for coherence), node records from L2.
// Limits: bodies whose positions agree in all 32 levels are merged like the reference's coincident
// bodies (`pos == existing_pos`, masses added in index order); the reference would subdivide
// further if their positions differ beyond that depth.
#include "kernels.h"
#include <cooperative_groups.h>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include "radix_sort.cuh"

namespace nb {

constexpr int BH_LEVELS = 32;

struct BhRoot { float cx, cy, size; int pad; };

// ---- 1. bounding box -------------------------------------------------------------------------
// floats map to unsigned keys whose integer order equals the float order, so min/max reduce with
// integer atomics; box[0..3] = min x, min y, max x, max y (encoded), reset by bh_reset_kernel.
__device__ __forceinline__ unsigned f2ord(float f)
{
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned k)
{
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

__global__ void bh_reset_kernel(unsigned *box, unsigned *count_tail)
{
    box[0] = box[1] = 0xffffffffu;   // running minima
    box[2] = box[3] = 0u;            // running maxima
    count_tail[0] = 0;               // count[n]   : terminates the exclusive scan
    count_tail[1] = 0;               // count[n+1] : deepest leaf level
}

__global__ void __launch_bounds__(256) bh_bbox_kernel(const float *__restrict__ posm, size_t n, unsigned *box)
{
    float minx = 3.402823466e+38f, miny = minx, maxx = -minx, maxy = -minx;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t g = blk_index(i, 0);
        const float x = posm[g], y = posm[g + BLK];
        minx = fminf(minx, x); maxx = fmaxf(maxx, x);
        miny = fminf(miny, y); maxy = fmaxf(maxy, y);
    }
    for (int o = 16; o > 0; o >>= 1) {
        minx = fminf(minx, __shfl_xor_sync(0xffffffffu, minx, o));
        miny = fminf(miny, __shfl_xor_sync(0xffffffffu, miny, o));
        maxx = fmaxf(maxx, __shfl_xor_sync(0xffffffffu, maxx, o));
        maxy = fmaxf(maxy, __shfl_xor_sync(0xffffffffu, maxy, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&box[0], f2ord(minx)); atomicMin(&box[1], f2ord(miny));
        atomicMax(&box[2], f2ord(maxx)); atomicMax(&box[3], f2ord(maxy));
    }
}

__global__ void bh_root_kernel(const unsigned *box, BhRoot *root)
{
    const float minx = ord2f(box[0]), miny = ord2f(box[1]), maxx = ord2f(box[2]), maxy = ord2f(box[3]);
    // Quad::new_containing, Quad.hpp:40-44: center = (min+max)*0.5f ; size = max(extent.x, extent.y)
    root->cx = __fmul_rn(__fadd_rn(minx, maxx), 0.5f);
    root->cy = __fmul_rn(__fadd_rn(miny, maxy), 0.5f);
    root->size = fmaxf(__fsub_rn(maxx, minx), __fsub_rn(maxy, miny));
}

// Quad::find_quadrant (Quad.hpp:47-49) and Quad::into_quadrant (Quad.hpp:51-57), one level down.
__device__ __forceinline__ unsigned bh_descend(float x, float y, float &cx, float &cy, float &size)
{
    const unsigned q = ((unsigned)(y > cy) << 1) | (unsigned)(x > cx);
    const float ns = __fmul_rn(size, 0.5f);
    cx = __fadd_rn(cx, __fmul_rn((q & 1u) ? 0.5f : -0.5f, ns));
    cy = __fadd_rn(cy, __fmul_rn((q & 2u) ? 0.5f : -0.5f, ns));
    size = ns;
    return q;
}

// ---- 2. quadrant-path keys -----------------------------------------------------------------------
__global__ void __launch_bounds__(256)
bh_keys_kernel(const float *__restrict__ posm, size_t n, const BhRoot *__restrict__ root,
               unsigned long long *__restrict__ keys, unsigned *__restrict__ idx)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const size_t g = blk_index(i, 0);
    const float x = posm[g], y = posm[g + BLK];
    float cx = root->cx, cy = root->cy, size = root->size;
    unsigned long long k = 0;
#pragma unroll 4
    for (int l = 0; l < BH_LEVELS; ++l) k = (k << 2) | bh_descend(x, y, cx, cy, size);
    keys[i] = k;
    idx[i] = (unsigned)i;
}

__device__ __forceinline__ int lcp_levels(unsigned long long a, unsigned long long b)
{
    return a == b ? BH_LEVELS : (__clzll((long long)(a ^ b)) >> 1);
}

// ---- 4. cells owned by each sorted body ------------------------------------------------------------
// first[s]..leaf[s] are the depths of the cells that first appear with sorted body s; duplicates
// (identical key as the previous body) own nothing.
__global__ void __launch_bounds__(256)
bh_count_kernel(const unsigned long long *__restrict__ keys, size_t n, unsigned *__restrict__ count,
                unsigned char *__restrict__ first, unsigned char *__restrict__ leaf, unsigned *__restrict__ max_depth)
{
    const size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const unsigned long long k = keys[s];
    if (s > 0 && keys[s - 1] == k) { count[s] = 0; first[s] = 0; leaf[s] = 0; return; }
    const int lp = (s > 0) ? lcp_levels(keys[s - 1], k) : -1;
    size_t t = s + 1;
    while (t < n && keys[t] == k) ++t;                       // skip the run of coincident bodies
    const int ln = (t < n) ? lcp_levels(k, keys[t]) : -1;
    const int leafd = (lp < 0 && ln < 0) ? 0 : max(lp, ln) + 1;   // a lone body is the root leaf
    const int firstd = lp + 1;                                // s == 0 -> 0: owns the root
    count[s] = (unsigned)(leafd - firstd + 1);
    first[s] = (unsigned char)firstd;
    leaf[s] = (unsigned char)leafd;
    // deepest leaf of the tree: bounds the number of propagate launches (warp-aggregated atomic)
    if ((unsigned)leafd > *max_depth) atomicMax(max_depth, (unsigned)leafd);   // racy pre-check only skips no-ops
}

// node record: com/body position, mass, size^2 ; next (0 = end of walk) ; depth | leaf flag
struct BhNodes {
    float4 *data;        // x, y, mass, size*size
    float4 *quad;        // cx, cy, size, unused (diagnostics / parity tests)
    unsigned *next;
    unsigned *meta;      // depth (low 8 bits) | leaf << 8
};

// ---- 5. emit the pre-order node array -------------------------------------------------------------
__global__ void __launch_bounds__(128)
bh_emit_kernel(const float *__restrict__ posm, const unsigned long long *__restrict__ keys,
               const unsigned *__restrict__ idx, size_t n, const BhRoot *__restrict__ root,
               const unsigned *__restrict__ offs, const unsigned *__restrict__ count,
               const unsigned char *__restrict__ first, const unsigned char *__restrict__ leaf,
               BhNodes nodes, unsigned cap)
{
    const size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n || count[s] == 0) return;
    const unsigned long long k = keys[s];
    const unsigned body = idx[s];
    const size_t g = blk_index(body, 0);
    const float x = posm[g], y = posm[g + BLK];
    // merged mass of coincident bodies, added in body-index order (stable sort) like insert() :56-60
    float mass = posm[g + 3 * BLK];
    for (size_t t = s + 1; t < n && keys[t] == k; ++t) mass = __fadd_rn(mass, posm[blk_index(idx[t], 0) + 3 * BLK]);
    const int firstd = first[s], leafd = leaf[s];
    const unsigned off = offs[s];
    float cx = root->cx, cy = root->cy, size = root->size;
    for (int d = 0; d <= leafd; ++d) {
        if (d >= firstd) {
            const unsigned c = off + (unsigned)(d - firstd);
            if (c < cap) {
                const bool is_leaf = (d == leafd);
                nodes.data[c] = make_float4(is_leaf ? x : 0.f, is_leaf ? y : 0.f, is_leaf ? mass : 0.f, __fmul_rn(size, size));
                nodes.quad[c] = make_float4(cx, cy, size, 0.f);
                nodes.meta[c] = (unsigned)d | (is_leaf ? 256u : 0u);
                // skip pointer: first sorted body after s whose depth-d prefix differs
                unsigned nx = 0;
                if (d > 0) {
                    const int sh = 2 * (BH_LEVELS - d);
                    const unsigned long long p = k >> sh;
                    size_t lo = s + 1, hi = n;           // first j in (s, n) with (keys[j] >> sh) > p
                    while (lo < hi) {
                        const size_t mid = (lo + hi) >> 1;
                        if ((keys[mid] >> sh) > p) hi = mid; else lo = mid + 1;
                    }
                    nx = (lo < n) ? offs[lo] : 0u;
                }
                nodes.next[c] = nx;
            }
        }
        if (d < leafd) {
            const unsigned q = (unsigned)(k >> (2 * (BH_LEVELS - 1 - d))) & 3u;
            const float ns = __fmul_rn(size, 0.5f);
            cx = __fadd_rn(cx, __fmul_rn((q & 1u) ? 0.5f : -0.5f, ns));
            cy = __fadd_rn(cy, __fmul_rn((q & 2u) ? 0.5f : -0.5f, ns));
            size = ns;
        }
    }
}

// ---- 6. centres of mass (Quadtree::propagate, :236-258) -------------------------------------------------
// One cooperative launch: levels deepest-first with a grid-wide barrier between levels.  Sorted body s
// owns the consecutive cells offs[s] .. offs[s]+count[s]-1 at depths first[s] .. leaf[s], so the branch
// cell of level L on its chain is found by index arithmetic -- no per-level node lists.
__device__ __forceinline__ void bh_propagate_cell(const BhNodes &nodes, unsigned c, unsigned m, unsigned level)
{
    const unsigned end = nodes.next[c];
    float px = 0.f, py = 0.f, mass = 0.f;
    unsigned ch = c + 1;                                   // children in quadrant order
    for (int i = 0; i < 4 && ch != end && ch < m; ++i) {
        if ((nodes.meta[ch] & 255u) != level + 1u) break;
        const float4 d = nodes.data[ch];
        px = __fadd_rn(px, __fmul_rn(d.x, d.z));
        py = __fadd_rn(py, __fmul_rn(d.y, d.z));
        mass = __fadd_rn(mass, d.z);
        const unsigned nx = nodes.next[ch];
        if (nx == 0) break;
        ch = nx;
    }
    if (mass > 0.f) { // Vec2::operator/=: inv = 1/scalar ; x *= inv ; y *= inv
        const float inv = __fdiv_rn(1.0f, mass);
        px = __fmul_rn(px, inv);
        py = __fmul_rn(py, inv);
    }
    float4 d = nodes.data[c];
    d.x = px; d.y = py; d.z = mass;
    nodes.data[c] = d;
}

__global__ void __launch_bounds__(256)
bh_propagate_kernel(BhNodes nodes, size_t n, const unsigned *__restrict__ offs, const unsigned char *__restrict__ first,
                    const unsigned char *__restrict__ leaf, const unsigned *__restrict__ count, unsigned cap)
{
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    const unsigned m = min(offs[n], cap);
    const int dmax = (int)min(count[n + 1], (unsigned)BH_LEVELS);
    for (int level = dmax - 1; level >= 0; --level) {
        for (size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x; s < n; s += (size_t)gridDim.x * blockDim.x) {
            const int f = first[s], l = leaf[s];
            if (count[s] != 0 && f <= level && level < l) {
                const unsigned c = offs[s] + (unsigned)(level - f);
                if (c < m) bh_propagate_cell(nodes, c, m, (unsigned)level);
            }
        }
        __threadfence();
        grid.sync();
    }
}

__device__ __forceinline__ float bh_quake(float number)
{
    const float y = __uint_as_float(0x5f3759dfu - (__float_as_uint(number) >> 1));
    return __fmul_rn(y, __fsub_rn(1.5f, __fmul_rn(__fmul_rn(__fmul_rn(number, 0.5f), y), y)));
}

// ---- 7. walk (Quadtree::acc, :113-155) ------------------------------------------------------------------
template <bool REFCOMPAT>
__global__ void __launch_bounds__(128)
bh_walk_kernel(const float *__restrict__ posm, const unsigned *__restrict__ idx, size_t n, BhNodes nodes,
               float t_sq, float e_sq, int fix_near_leaves, size_t shard_start, size_t shard_count,
               float *__restrict__ accp, unsigned cap)
{
    const size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const unsigned body = idx[s];                           // targets in Z-order: neighbouring threads walk alike
    if (body < shard_start || body >= shard_start + shard_count) return;
    const size_t g = blk_index(body, 0);
    const float px = posm[g], py = posm[g + BLK];
    float ax = 0.f, ay = 0.f;
    unsigned i = 0;
    do {
        const float4 nd = nodes.data[i];
        const float dx = __fsub_rn(nd.x, px), dy = __fsub_rn(nd.y, py);
        const float d_sq = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
        const bool far = nd.w < __fmul_rn(d_sq, t_sq);
        const bool is_leaf = (nodes.meta[i] & 256u) != 0u;
        if (far || is_leaf) {
            if ((far || fix_near_leaves) && d_sq > 0.f) {
                float s3;
                if (REFCOMPAT) {
                    const float inv = bh_quake(__fadd_rn(d_sq, e_sq));
                    s3 = __fmul_rn(nd.z, __fmul_rn(__fmul_rn(inv, inv), inv));
                } else {
                    const float inv = rsqrt_approx(d_sq + e_sq);
                    s3 = nd.z * inv * inv * inv;
                }
                ax = __fadd_rn(ax, __fmul_rn(dx, s3));
                ay = __fadd_rn(ay, __fmul_rn(dy, s3));
            }
            i = nodes.next[i];
        } else {
            i = i + 1;
        }
    } while (i != 0 && i < cap);   // i >= cap only if the tree overflowed its reservation (reported by node_count)
    const size_t l = blk_index(body - shard_start, 0);
    accp[l] = ax; accp[l + BLK] = ay; accp[l + 2 * BLK] = 0.f;
}

// Warp-cooperative walk.  The 32 lanes of a warp hold 32 targets that are neighbours in Z-order; the
// warp walks the UNION of their traversals in pre-order index order and every visited node record is
// loaded once per warp (uniform address -> one broadcast transaction instead of up to 32 divergent
// ones).  Each lane keeps `resume`, the index of the next node of ITS OWN sequential walk (node+1 when
// it opens a branch, next[node] when it accepts a far node or passes a leaf); it takes part only at
// nodes == resume, and the warp advances to the minimum resume over lanes.  A lane's sequence of
// visited nodes, and therefore its sum term by term, is exactly that of Quadtree::acc -- bit-exact
// like bh_walk_kernel, only the memory access pattern differs.
template <bool REFCOMPAT>
__global__ void __launch_bounds__(128)
bh_walk_warp_kernel(const float *__restrict__ posm, const unsigned *__restrict__ idx, size_t n, BhNodes nodes,
                    float t_sq, float e_sq, int fix_near_leaves, size_t shard_start, size_t shard_count,
                    float *__restrict__ accp, unsigned cap)
{
    const size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = s < n;
    const unsigned body = valid ? idx[s] : 0u;
    const bool mine = valid && body >= shard_start && body < shard_start + shard_count;
    const size_t g = blk_index(body, 0);
    const float px = posm[g], py = posm[g + BLK];
    float ax = 0.f, ay = 0.f;
    constexpr unsigned DONE = 0xffffffffu;
    unsigned resume = mine ? 0u : DONE;
    unsigned i = __reduce_min_sync(0xffffffffu, resume);
    while (i < cap) {                                        // DONE (and an overflowed tree) end the walk
        const float4 nd = nodes.data[i];                     // warp-uniform loads
        const unsigned nx = nodes.next[i];
        const bool is_leaf = (nodes.meta[i] & 256u) != 0u;
        if (resume == i) {
            const float dx = __fsub_rn(nd.x, px), dy = __fsub_rn(nd.y, py);
            const float d_sq = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
            const bool far = nd.w < __fmul_rn(d_sq, t_sq);
            if (far || is_leaf) {
                if ((far || fix_near_leaves) && d_sq > 0.f) {
                    float s3;
                    if (REFCOMPAT) {
                        const float inv = bh_quake(__fadd_rn(d_sq, e_sq));
                        s3 = __fmul_rn(nd.z, __fmul_rn(__fmul_rn(inv, inv), inv));
                    } else {
                        const float inv = rsqrt_approx(d_sq + e_sq);
                        s3 = nd.z * inv * inv * inv;
                    }
                    ax = __fadd_rn(ax, __fmul_rn(dx, s3));
                    ay = __fadd_rn(ay, __fmul_rn(dy, s3));
                }
                resume = nx ? nx : DONE;
            } else {
                resume = i + 1;
            }
        }
        i = __reduce_min_sync(0xffffffffu, resume);
    }
    if (mine) {
        const size_t l = blk_index(body - shard_start, 0);
        accp[l] = ax; accp[l + BLK] = ay; accp[l + 2 * BLK] = 0.f;
    }
}

// ---- host side ------------------------------------------------------------------------------------------------
cudaError_t BhWorkspace::alloc(size_t n)
{
    cudaError_t e;
    n_cap = n;
    node_cap = (unsigned)std::min<size_t>(4 * n + 1024, 0x7fffffffu);
#define BH_ALLOC(p, bytes) if ((e = cudaMalloc((void **)&(p), (bytes))) != cudaSuccess) return e;
    BH_ALLOC(root, sizeof(BhRoot)) BH_ALLOC(box, 16)
    BH_ALLOC(keys_in, n * 8) BH_ALLOC(keys, n * 8) BH_ALLOC(idx_in, n * 4) BH_ALLOC(idx, n * 4)
    BH_ALLOC(count, (n + 2) * 4) BH_ALLOC(offs, (n + 2) * 4) BH_ALLOC(first, n) BH_ALLOC(leaf, n)
    BH_ALLOC(node_data, (size_t)node_cap * 16) BH_ALLOC(node_quad, (size_t)node_cap * 16)
    BH_ALLOC(node_next, (size_t)node_cap * 4) BH_ALLOC(node_meta, (size_t)node_cap * 4)
    size_t t1 = 0, t2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, t1, (unsigned long long *)nullptr, (unsigned long long *)nullptr,
                                    (unsigned *)nullptr, (unsigned *)nullptr, (int)n, 0, 64);
    cub::DeviceScan::ExclusiveSum(nullptr, t2, (unsigned *)nullptr, (unsigned *)nullptr, (int)n + 1);
    temp_bytes = std::max(std::max(t1, t2), radix_sort_temp_bytes(n));
    BH_ALLOC(temp, temp_bytes)
    {   // co-resident grid size for the cooperative propagate kernel
        int dev = 0, sms = 0, per_sm = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bh_propagate_kernel, 256, 0);
        coop_blocks = std::max(1, sms * std::max(1, std::min(per_sm, 4)));
    }
#undef BH_ALLOC
    return cudaSuccess;
}

void BhWorkspace::release()
{
    void *ptrs[] = {root, box, keys_in, keys, idx_in, idx, count, offs, first, leaf, node_data, node_quad, node_next, node_meta, temp};
    for (void *p : ptrs) if (p) cudaFree(p);
    *this = BhWorkspace();
}

static BhNodes bh_nodes(const BhWorkspace &w)
{
    BhNodes nd;
    nd.data = (float4 *)w.node_data; nd.quad = (float4 *)w.node_quad;
    nd.next = (unsigned *)w.node_next; nd.meta = (unsigned *)w.node_meta;
    return nd;
}

// Build the tree of the first n bodies of `posm`.  Fully asynchronous (no host read-back): the node count
// stays on the device (offs[n]); n_nodes is fetched lazily by node_count().
cudaError_t BhWorkspace::build(const float *posm, size_t n, cudaStream_t st, int *launches)
{
    if (n == 0 || n > n_cap) return cudaErrorInvalidValue;
    cudaError_t e;
    const unsigned g256 = (unsigned)((n + 255) / 256), g128 = (unsigned)((n + 127) / 128);
    unsigned *cnt = (unsigned *)count;
    bh_reset_kernel<<<1, 1, 0, st>>>((unsigned *)box, cnt + n);
    bh_bbox_kernel<<<std::min(g256, 4u * 148u), 256, 0, st>>>(posm, n, (unsigned *)box);
    bh_root_kernel<<<1, 1, 0, st>>>((const unsigned *)box, (BhRoot *)root);
    bh_keys_kernel<<<g256, 256, 0, st>>>(posm, n, (const BhRoot *)root, (unsigned long long *)keys_in, (unsigned *)idx_in);
    size_t tb = temp_bytes;
    if (own_sort) { // stable LSD sort; the result lands back in the first buffer pair -> swap roles
        if ((e = radix_sort_u64((unsigned long long *)keys_in, (unsigned long long *)keys, (unsigned *)idx_in, (unsigned *)idx,
                                n, temp, st, 0, 64, launches)) != cudaSuccess) return e;
        std::swap(keys_in, keys);
        std::swap(idx_in, idx);
    } else if ((e = cub::DeviceRadixSort::SortPairs(temp, tb, (const unsigned long long *)keys_in, (unsigned long long *)keys,
                                                    (const unsigned *)idx_in, (unsigned *)idx, (int)n, 0, 64, st)) != cudaSuccess) return e;
    bh_count_kernel<<<g256, 256, 0, st>>>((const unsigned long long *)keys, n, cnt, (unsigned char *)first,
                                          (unsigned char *)leaf, cnt + n + 1);
    tb = temp_bytes;
    if ((e = cub::DeviceScan::ExclusiveSum(temp, tb, (const unsigned *)count, (unsigned *)offs, (int)n + 1, st)) != cudaSuccess) return e;
    bh_emit_kernel<<<g128, 128, 0, st>>>(posm, (const unsigned long long *)keys, (const unsigned *)idx, n, (const BhRoot *)root,
                                         (const unsigned *)offs, (const unsigned *)count, (const unsigned char *)first,
                                         (const unsigned char *)leaf, bh_nodes(*this), node_cap);
    {
        BhNodes nd = bh_nodes(*this);
        size_t nn = n;
        const unsigned *po = (const unsigned *)offs, *pc = (const unsigned *)count;
        const unsigned char *pf = (const unsigned char *)first, *pl = (const unsigned char *)leaf;
        unsigned cap = node_cap;
        void *args[] = {&nd, &nn, &po, &pf, &pl, &pc, &cap};
        const unsigned grid = std::min(g256, (unsigned)coop_blocks);
        if ((e = cudaLaunchCooperativeKernel((void *)bh_propagate_kernel, dim3(grid), dim3(256), args, 0, st)) != cudaSuccess) return e;
    }
    count_valid = false;
    if (launches) *launches += 7 + 3;                        // own kernels + the sort/scan passes (counted as 3)
    return cudaGetLastError();
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

cudaError_t BhWorkspace::walk(const float *posm, size_t n, float theta, float eps, bool refcompat, bool fix_near_leaves,
                              size_t shard_start, size_t shard_count, float *accp, cudaStream_t st)
{
    const unsigned g = (unsigned)((n + 127) / 128);
    const float t_sq = theta * theta, e_sq = eps * eps;       // Quadtree ctor, Quadtree.hpp:19
    if (warp_walk) {
        if (refcompat)
            bh_walk_warp_kernel<true><<<g, 128, 0, st>>>(posm, (const unsigned *)idx, n, bh_nodes(*this), t_sq, e_sq,
                                                         fix_near_leaves ? 1 : 0, shard_start, shard_count, accp, node_cap);
        else
            bh_walk_warp_kernel<false><<<g, 128, 0, st>>>(posm, (const unsigned *)idx, n, bh_nodes(*this), t_sq, e_sq,
                                                          fix_near_leaves ? 1 : 0, shard_start, shard_count, accp, node_cap);
        return cudaGetLastError();
    }
    if (refcompat)
        bh_walk_kernel<true><<<g, 128, 0, st>>>(posm, (const unsigned *)idx, n, bh_nodes(*this), t_sq, e_sq,
                                                fix_near_leaves ? 1 : 0, shard_start, shard_count, accp, node_cap);
    else
        bh_walk_kernel<false><<<g, 128, 0, st>>>(posm, (const unsigned *)idx, n, bh_nodes(*this), t_sq, e_sq,
                                                 fix_near_leaves ? 1 : 0, shard_start, shard_count, accp, node_cap);
    return cudaGetLastError();
}

// node array in walk order for the parity tests: f6 = (x, y, mass, cx, cy, size), u2 = (next, depth | leaf<<8)
cudaError_t BhWorkspace::download_nodes(float *f6, unsigned *u2, size_t cap, cudaStream_t st)
{
    const size_t m = std::min<size_t>(std::min<size_t>(cap, n_nodes), node_cap);
    std::vector<float4> d(m), q(m);
    std::vector<unsigned> nx(m), me(m);
    cudaError_t e;
    if ((e = cudaMemcpyAsync(d.data(), node_data, m * 16, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
    if ((e = cudaMemcpyAsync(q.data(), node_quad, m * 16, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
    if ((e = cudaMemcpyAsync(nx.data(), node_next, m * 4, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
    if ((e = cudaMemcpyAsync(me.data(), node_meta, m * 4, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;
    for (size_t i = 0; i < m; ++i) {
        f6[6 * i + 0] = d[i].x; f6[6 * i + 1] = d[i].y; f6[6 * i + 2] = d[i].z;
        f6[6 * i + 3] = q[i].x; f6[6 * i + 4] = q[i].y; f6[6 * i + 5] = q[i].z;
        u2[2 * i + 0] = nx[i]; u2[2 * i + 1] = me[i];
    }
    return cudaSuccess;
}

} // namespace nb
