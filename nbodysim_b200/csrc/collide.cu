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
// What can be dropped, and what runs in parallel.  resolve() (:293-346) moves a body only when the pair
// overlaps at the moment it is resolved, so a sweep pair can matter only if one of its bodies may have been
// moved before its turn -- i.e. if it is connected, through sweep pairs, to a pair that overlaps NOW.  The
// sweep pairs form a graph on the bodies; its connected components are found with a lock-free union-find
// (hooking the larger root under the smaller, so a component's label is its smallest body index);
// components that contain an overlapping pair are "hot".  Exactly the pairs of hot components are kept
// (dropping the others cannot change the result: none of their resolves would pass the overlap test), and
// because two components share no body, components are resolved IN PARALLEL -- one thread per component,
// walking its pairs in the canonical (first, second) order, each resolve reading what the previous one
// wrote, as in the reference.  Pair keys sort as (component, first, second).
//
// The phases are written as grid-stride device functions so that the same code runs either as one kernel
// per phase (any n) or inside the single-cluster kernel of small scenes (cluster barriers between phases).
#include "collide.cuh"
#include "cluster_prims.cuh"
#include <cstring>

namespace nb {

__global__ void __launch_bounds__(256) col_init_kernel(ColArgs a)
{
    col_phase_init(a, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}
__global__ void __launch_bounds__(256) col_entries_kernel(ColArgs a)
{
    col_phase_entries(a, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}
__global__ void __launch_bounds__(256) col_detect_kernel(ColArgs a)
{
    col_phase_pairs<0>(a, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}
__global__ void __launch_bounds__(256) col_union_kernel(ColArgs a)
{
    col_phase_pairs<1>(a, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}
__global__ void __launch_bounds__(256) col_mark_kernel(ColArgs a)
{
    col_phase_mark(a, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}
__global__ void __launch_bounds__(256) col_emit_kernel(ColArgs a)
{
    col_phase_pairs<2>(a, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}
__global__ void __launch_bounds__(128) col_resolve_kernel(ColArgs a, const unsigned long long *__restrict__ pairs_sorted)
{
    col_phase_resolve(a, pairs_sorted, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}

// ---- screening kernels (every pass) -----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) col_grid_insert_kernel(ColArgs a, ColGrid g)
{
    col_grid_insert(a, g, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}
__global__ void __launch_bounds__(256) col_grid_detect_kernel(ColArgs a, ColGrid g)
{
    col_grid_detect(a, g, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}

// ---- the full pass as ONE CTA (small scenes) ----------------------------------------------------------------------------------
// Launched after the screening of every pass, it returns at once unless the screening saw an overlap: a step without
// collisions costs a memset, two short kernels and this empty launch.  When it does run, the phases of the full pass
// follow one another separated by CTA-wide barriers (the cluster primitives with a cluster of one CTA; sort ranks in a
// global scratch array).  Collisions are rare events in the scenes this path serves, so its own speed matters little.
__global__ void __launch_bounds__(CL_THREADS, 1)
col_single_cta_kernel(ColArgs a, unsigned long long *keys_b, unsigned *vals_b, unsigned long long *pairs_b, unsigned *ranks, int key_bits)
{
    __shared__ ClSmem sm;
    const unsigned gtid = threadIdx.x, gthreads = CL_THREADS;
    if (gtid < 8) a.counters[gtid] = 0;                                   // the statistics of a pass that found nothing
    if (col_gate_closed(a)) return;
    a.gate = nullptr;
    cl_sync();
    col_phase_init(a, gtid, gthreads);
    cl_sync();
    col_phase_entries(a, gtid, gthreads);
    cl_sync();
    const unsigned ne = min(__ldcg(a.counters + 0), a.entry_cap);
    if (__ldcg(a.counters + 2)) return;                                   // uniform: every thread reads the same word
    if (cl_radix_sort<true>(sm, ranks, a.keys_in, keys_b, a.vals_in, vals_b, ne, 0, COL_GROUP_BITS)) { a.keys = keys_b; a.vals = vals_b; }
    else { a.keys = a.keys_in; a.vals = a.vals_in; }
    col_phase_pairs<0>(a, gtid, gthreads);
    cl_sync();
    if (__ldcg(a.counters + 4) == 0) return;                              // no SWEEP pair overlaps: no resolve could pass its test
    col_phase_pairs<1>(a, gtid, gthreads);
    cl_sync();
    col_phase_mark(a, gtid, gthreads);
    cl_sync();
    col_phase_pairs<2>(a, gtid, gthreads);
    cl_sync();
    const unsigned np = __ldcg(a.counters + 1);
    if (__ldcg(a.counters + 2) || np > a.pair_cap) return;
    const unsigned long long *sorted = cl_radix_sort<false>(sm, ranks, a.pairs, pairs_b, nullptr, nullptr, np, 0, key_bits) ? pairs_b : a.pairs;
    col_phase_resolve(a, sorted, gtid, gthreads);
}

// single_cta_mode: 0 = scenes of up to single_cta_max_n bodies run their (rare) full pass as one CTA, 1 = never (one
// kernel per phase), 2 = always
cudaError_t CollideWorkspace::alloc(size_t n, int single_cta_mode, size_t single_cta_max_n)
{
    cudaError_t e;
    single_cta = single_cta_mode == 2 || (single_cta_mode == 0 && n <= single_cta_max_n);
    entry_cap = (unsigned)std::min<size_t>(4 * n + 4096, 0x7fffffffu);
    pair_cap = (unsigned)std::min<size_t>(64 * n + 65536, (size_t)16 << 20);   // pairs of hot components
    unsigned tsize = 1024;
    while (tsize < 2 * (size_t)entry_cap && tsize < (1u << 30)) tsize <<= 1;   // open addressing at load <= 1/2
    table_slots = tsize;
#define COL_ALLOC(p, bytes) if ((e = cudaMalloc((void **)&(p), (bytes))) != cudaSuccess) return e;
    COL_ALLOC(keys_in, (size_t)entry_cap * 8) COL_ALLOC(keys, (size_t)entry_cap * 8)
    COL_ALLOC(vals_in, (size_t)entry_cap * 4) COL_ALLOC(vals, (size_t)entry_cap * 4)
    COL_ALLOC(pairs_in, (size_t)pair_cap * 8) COL_ALLOC(pairs, (size_t)pair_cap * 8)
    COL_ALLOC(hot, 2 * n) COL_ALLOC(parent, n * 4) COL_ALLOC(counters, 32)
    // screening: flags (64 B) | table keys | list heads -- one region, one memset per pass; then the list links
    grid_bytes = 64 + (size_t)tsize * 12;
    COL_ALLOC(grid, grid_bytes)
    COL_ALLOC(grid_links, (size_t)entry_cap * 8)
    if (single_cta) { COL_ALLOC(ranks, (size_t)std::max(entry_cap, pair_cap) * 4) }
    temp_bytes = radix_sort_temp_bytes(std::max<size_t>(entry_cap, pair_cap));
    COL_ALLOC(temp, temp_bytes)
#undef COL_ALLOC
    n_cap = n;
    return cudaSuccess;
}

void CollideWorkspace::release()
{
    void *ptrs[] = {keys_in, keys, vals_in, vals, pairs_in, pairs, hot, parent, counters, temp, grid, grid_links, ranks};
    for (void *p : ptrs) if (p) cudaFree(p);
    *this = CollideWorkspace();
}

ColArgs CollideWorkspace::args(float *posm, float *vel, size_t n) const
{
    ColArgs a;
    a.posm = posm; a.vel = vel; a.n = (unsigned)n;
    a.keys_in = (unsigned long long *)keys_in; a.vals_in = (unsigned *)vals_in;
    a.keys = (const unsigned long long *)keys; a.vals = (const unsigned *)vals;
    a.entry_cap = entry_cap;
    a.pairs = (unsigned long long *)pairs_in; a.pair_cap = pair_cap;
    a.hot = (unsigned char *)hot; a.parent = (unsigned *)parent; a.counters = (unsigned *)counters;
    a.status = status;
    a.gate = (const unsigned *)grid;       // flags[0] of the screening
    int idx_bits = 1;
    while (((size_t)1 << idx_bits) < n) ++idx_bits;
    a.idx_bits = idx_bits;
    a.rooted = (3 * idx_bits <= 64) ? 1 : 0;     // beyond 2^21 bodies the component label no longer fits the key: one serial run
    return a;
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
    ColArgs a = args(posm, vel, n);
    const int key_bits = std::min(64, (((a.rooted ? 3 : 2) * a.idx_bits + 7) / 8) * 8);
    {   // screening: hash-grid insert + overlap test; everything after it is gated on its flag
        ColGrid g;
        g.flags = (unsigned *)grid;
        g.tkeys = (unsigned long long *)((char *)grid + 64);
        g.heads = (unsigned *)((char *)grid + 64 + (size_t)table_slots * 8);
        g.enext = (unsigned *)grid_links; g.ebody = (unsigned *)grid_links + entry_cap;
        g.tmask = table_slots - 1; g.ecap = entry_cap;
        if ((e = cudaMemsetAsync(grid, 0, grid_bytes, st)) != cudaSuccess) return e;
        const unsigned gb = (unsigned)((n + 255) / 256);
        col_grid_insert_kernel<<<gb, 256, 0, st>>>(a, g);
        col_grid_detect_kernel<<<gb, 256, 0, st>>>(a, g);
        if (launches) *launches += 2;
    }
    if (single_cta) {
        // an explicit cluster of ONE CTA: the phase barriers are the cluster primitives of cluster_prims.cuh
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof cfg);
        cfg.gridDim = dim3(1); cfg.blockDim = dim3(CL_THREADS); cfg.stream = st;
        cudaLaunchAttribute at;
        at.id = cudaLaunchAttributeClusterDimension;
        at.val.clusterDim.x = 1; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
        cfg.attrs = &at; cfg.numAttrs = 1;
        if ((e = cudaLaunchKernelEx(&cfg, col_single_cta_kernel, a, (unsigned long long *)keys, (unsigned *)vals, (unsigned long long *)pairs,
                                    (unsigned *)ranks, key_bits)) != cudaSuccess) return e;
        if (launches) *launches += 1;
        return cudaGetLastError();
    }
    const unsigned gn = (unsigned)((n + 255) / 256), ge = (entry_cap + 255) / 256;
    col_init_kernel<<<gn, 256, 0, st>>>(a);
    col_entries_kernel<<<gn, 256, 0, st>>>(a);
    // only the GROUPING of the entries matters to the pair discovery: sort on COL_GROUP_BITS bits of the hash (collide.cuh)
    if ((e = radix_sort_u64((unsigned long long *)keys_in, (unsigned long long *)keys, (unsigned *)vals_in, (unsigned *)vals,
                            entry_cap, temp, st, 0, COL_GROUP_BITS, launches, cnt + 0)) != cudaSuccess) return e;
    std::swap(keys_in, keys);
    std::swap(vals_in, vals);
    a = args(posm, vel, n);
    col_detect_kernel<<<ge, 256, 0, st>>>(a);
    col_union_kernel<<<ge, 256, 0, st>>>(a);     // this and the next two exit at once when nothing overlaps
    col_mark_kernel<<<gn, 256, 0, st>>>(a);
    col_emit_kernel<<<ge, 256, 0, st>>>(a);
    // a pair key packs (component, first, second) into 3 x idx_bits bits: 6 digit passes at n = 25,000
    if ((e = radix_sort_u64((unsigned long long *)pairs_in, (unsigned long long *)pairs, nullptr, nullptr, pair_cap, temp, st,
                            0, key_bits, launches, cnt + 1)) != cudaSuccess) return e;
    std::swap(pairs_in, pairs);
    col_resolve_kernel<<<std::min((pair_cap + 127) / 128, 148u * 8u), 128, 0, st>>>(a, (const unsigned long long *)pairs);
    if (launches) *launches += 7;
    return cudaGetLastError();
}

// counters of the last pass: [0] cell entries, [1] pairs kept, [2] overflow flag, [3] pairs resolved (synchronises)
cudaError_t CollideWorkspace::stats(cudaStream_t st, unsigned out[4])
{
    cudaError_t e;
    if ((e = cudaMemcpyAsync(out, counters, 16, cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
    return cudaStreamSynchronize(st);
}

} // namespace nb
