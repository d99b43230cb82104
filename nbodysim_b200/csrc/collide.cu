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
// Every pass starts with a hash grid over the bodies' cells (collide.cuh).  Small scenes take their sweep pairs
// straight from it and finish in one CTA; large scenes use it to decide whether anything overlaps at all before
// running one kernel per phase.  The phases are grid-stride device functions shared by both.
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

__global__ void __launch_bounds__(256) col_grid_pairs_kernel(ColArgs a, ColGrid g)
{
    pdl_enter();
    col_grid_pairs(a, g, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}

// ---- small scenes: the rest of the pass as ONE CTA ------------------------------------------------------------------------------
// The shipped scene has a few hundred sweep pairs among its 25,000 bodies and, once bodies start falling into the
// central mass, a few dozen collisions per step: too little work to spread over launches.  One CTA takes the pair list
// of col_grid_pairs and, unless nothing overlaps, runs the remaining phases separated by CTA barriers: union-find over
// the pairs, hot components, the kept pairs keyed (component, first, second), their sort (cluster_prims.cuh with a
// cluster of one CTA), and the resolve -- one thread per component, pairs in canonical order.
__global__ void __launch_bounds__(CL_THREADS, 1)
col_finish_kernel(ColArgs a, ColGrid g, unsigned long long *pairs_b, unsigned *ranks, int key_bits)
{
    __shared__ ClSmem sm;
    pdl_enter();
    const unsigned tid = threadIdx.x;
    if (tid < 8) a.counters[tid] = 0;                                     // the statistics of a pass that found nothing
    const unsigned P = min(__ldcg(g.flags + 2), a.pair_cap);
    if (__ldcg(g.flags + 0) == 0u || P == 0u || __ldcg(g.flags + 2) > a.pair_cap) return;   // uniform
    __syncthreads();
    const int b = a.idx_bits;
    const unsigned long long imask = (1ull << b) - 1ull;
    for (unsigned p = tid; p < P; p += CL_THREADS) {                      // the bodies that occur in pairs: singleton sets, cold
        const unsigned long long k = a.pairs[p];
        const unsigned f = (unsigned)((k >> b) & imask), s2 = (unsigned)(k & imask);
        a.parent[f] = f; a.parent[s2] = s2;
        a.hot[a.n + f] = 0; a.hot[a.n + s2] = 0;
    }
    __syncthreads();
    for (unsigned p = tid; p < P; p += CL_THREADS) {
        const unsigned long long k = a.pairs[p];
        uf_union(a.parent, (unsigned)((k >> b) & imask), (unsigned)(k & imask));
    }
    __syncthreads();
    for (unsigned p = tid; p < P; p += CL_THREADS) {
        const unsigned long long k = a.pairs[p];
        if (k & COL_OVERLAP_BIT) a.hot[a.n + uf_find(a.parent, (unsigned)((k >> b) & imask))] = 1;
    }
    __syncthreads();
    for (unsigned p = tid; p < P; p += CL_THREADS) {
        const unsigned long long k = a.pairs[p] & ~COL_OVERLAP_BIT;
        const unsigned root = uf_find(a.parent, (unsigned)((k >> b) & imask));
        if (a.hot[a.n + root]) {
            const unsigned q = atomicAdd(&a.counters[1], 1u);
            pairs_b[q] = a.rooted ? (k | ((unsigned long long)root << (2 * b))) : k;
        }
    }
    if (tid == 0) { a.counters[4] = __ldcg(g.flags + 0); a.counters[0] = __ldcg(g.flags + 1); }
    __syncthreads();
    const unsigned np = __ldcg(a.counters + 1);
    const unsigned long long *sorted;
    constexpr unsigned RANK_SORT_MAX = sizeof(sm.counts) / sizeof(unsigned long long);   // 2048 keys fit the counter array
    if (np <= RANK_SORT_MAX) {
        // a handful of pairs: rank sort in shared memory (every key counts the keys before it; ties -- the same pair met in
        // several shared cells -- keep their order), one barrier instead of 6 radix digits
        unsigned long long *sk = reinterpret_cast<unsigned long long *>(&sm.counts[0][0]);
        for (unsigned p = tid; p < np; p += CL_THREADS) sk[p] = pairs_b[p];
        __syncthreads();
        for (unsigned p = tid; p < np; p += CL_THREADS) {
            const unsigned long long k = sk[p];
            unsigned r = 0;
            for (unsigned q = 0; q < np; ++q) { const unsigned long long o = sk[q]; r += (o < k || (o == k && q < p)) ? 1u : 0u; }
            a.pairs[r] = k;
        }
        __syncthreads();
        sorted = a.pairs;
    } else {
        sorted = cl_radix_sort<false>(sm, ranks, pairs_b, a.pairs, nullptr, nullptr, np, 0, key_bits) ? a.pairs : pairs_b;
    }
    col_phase_resolve(a, sorted, tid, CL_THREADS);
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
    COL_ALLOC(grid_links, (size_t)entry_cap * 20)
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
    static const float strip_env = getenv("NBODY_COL_STRIP") ? (float)atof(getenv("NBODY_COL_STRIP")) : COL_STRIP;   // tuning
    a.strip = strip_env;
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
ColGrid CollideWorkspace::grid_view() const
{
    ColGrid g;
    g.flags = (unsigned *)grid;
    g.tkeys = (unsigned long long *)((char *)grid + 64);
    g.heads = (unsigned *)((char *)grid + 64 + (size_t)table_slots * 8);
    g.edata = (float4 *)grid_links; g.enext = (unsigned *)((char *)grid_links + (size_t)entry_cap * 16);
    g.tmask = table_slots - 1; g.ecap = entry_cap;
    return g;
}

cudaError_t CollideWorkspace::prepare(cudaStream_t st) { return cudaMemsetAsync(grid, 0, grid_bytes, st); }

cudaError_t CollideWorkspace::run(float *posm, float *vel, size_t n, cudaStream_t st, int *launches, bool grid_filled)
{
    if (n == 0 || n > n_cap) return cudaErrorInvalidValue;
    cudaError_t e;
    unsigned *cnt = (unsigned *)counters;
    ColArgs a = args(posm, vel, n);
    const int key_bits = std::min(64, (((a.rooted ? 3 : 2) * a.idx_bits + 7) / 8) * 8);
    // the hash grid over the bodies' cells; what follows it is gated on whether anything overlaps
    const ColGrid g = grid_view();
    const unsigned gb = (unsigned)((n + 255) / 256);
    if (!grid_filled) {
        if ((e = prepare(st)) != cudaSuccess) return e;
        col_grid_insert_kernel<<<gb, 256, 0, st>>>(a, g);
        if (launches) *launches += 1;
    }
    if (single_cta) {   // small scene: sweep pairs straight from the grid, everything else in one CTA
        if ((e = launch_pdl(col_grid_pairs_kernel, dim3(gb), dim3(256), 0, st, a, g)) != cudaSuccess) return e;
        // an explicit cluster of ONE CTA: its sort uses the cluster primitives of cluster_prims.cuh
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof cfg);
        cfg.gridDim = dim3(1); cfg.blockDim = dim3(CL_THREADS); cfg.stream = st;
        cudaLaunchAttribute at[2];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;    // may become resident while the pair kernel runs (pdl_enter)
        at[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 2 : 1;
        if ((e = cudaLaunchKernelEx(&cfg, col_finish_kernel, a, g, (unsigned long long *)pairs, (unsigned *)ranks, key_bits)) != cudaSuccess) return e;
        if (launches) *launches += 2;
        return cudaGetLastError();
    }
    col_grid_detect_kernel<<<gb, 256, 0, st>>>(a, g);
    if (launches) *launches += 1;
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
