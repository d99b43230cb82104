// common.cuh -- layout constants, error plumbing and sm_100a PTX helpers shared by the kernels.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace nb {

// ---------------------------------------------------------------------------------------------
// Device body storage: "blocked SoA".  Bodies are grouped in blocks of BLK = 256; one block is
//     T x[256], y[256], z[256], m[256]          (4 KiB for float, 8 KiB for double)
// so (a) a thread reading lane t of a component is perfectly coalesced, (b) a source tile of any
// whole number of blocks is ONE contiguous range -> one cp.async.bulk (TMA 1-D bulk copy) stages
// it, and lands in shared memory already in SoA form, where a single LDS.128 yields four
// consecutive x (or y, z, m) values = two packed f32x2 operands, and (c) a rank's shard is a
// contiguous byte range -> in-place ncclAllGather.  vel uses the same blocks with (vx,vy,vz,radius),
// acc with (ax,ay,az,unused).
// Replaces the reference's 64-byte AoS Body (Body.hpp:6-14), of which the force loop touches 12 B.
// ---------------------------------------------------------------------------------------------
constexpr int BLK = 256;                 // bodies per block
constexpr int BLK_ELEMS = 4 * BLK;       // scalars per block
constexpr float PAD_POS = 1.0e18f;       // coordinates of the zero-mass padding bodies

__host__ __device__ inline size_t blk_index(size_t body, int comp)
{
    return (body >> 8) * (size_t)BLK_ELEMS + (size_t)comp * BLK + (body & 255);
}

// ---- programmatic dependent launch (griddepcontrol) -------------------------------------------------------------------------
// The Barnes-Hut step is ten short kernels in a row; a kernel launched with launch_pdl() may become RESIDENT while its
// predecessor is still running and waits in pdl_enter() until that one has completed (and its writes are visible), so the
// ~1.5 us a kernel boundary costs inside a graph overlaps with the predecessor's tail.  pdl_enter() is the first statement
// of every kernel on that path: wait for the predecessor, THEN allow the successor to be made resident (at most one kernel
// ahead).  In a kernel that was launched normally both instructions do nothing.
__device__ __forceinline__ void pdl_enter()
{
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;");
}
// NBODY_PDL=0: plain launches
inline bool pdl_enabled()
{
    static const bool on = !(getenv("NBODY_PDL") && atoi(getenv("NBODY_PDL")) == 0);
    return on;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args)
{
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at;
    at.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- sm_100a PTX helpers: mbarrier + 1-D TMA bulk copy (cp.async.bulk -> SASS UBLKCP) ----------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra.uni WAIT_DONE;\n"
        "bra.uni WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy; completion is signalled as transaction bytes on `bar`.
__device__ __forceinline__ void tma_bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes,
                                             uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
// ---- cross-process exchange: completion flags in peer memory (CUDA IPC over NVLink) -------------
// Every rank owns an array `flags[world]` of 64-bit step counters in its own HBM; flags[r] is written only by
// rank r (through an IPC mapping) and says "all positions rank r pushed for steps < flags[r] have landed here".
struct PeerSignal {
    unsigned long long *slot[16];   // slot[k]: this rank's counter inside peer k's flag array
    int n;                          // peers to signal (0: no signalling, e.g. single GPU)
    unsigned long long value;       // number of integrate-and-push steps completed once this launch is done
    unsigned *arrive;               // CTA arrival counter of this launch (device-local, self-resetting)
};
__device__ __forceinline__ void st_release_sys_u64(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// Called by EVERY thread of the grid after its last peer store.  The last CTA to arrive publishes `value` to
// every peer: each thread fences its own stores at system scope, the CTA's arrival is a device-scope atomic,
// and the publishing thread fences again before the release stores (fence-atomic / atomic-fence chains make
// all CTAs' stores happen-before the flag).
__device__ __forceinline__ void signal_peers_when_grid_done(const PeerSignal &sig)
{
    if (sig.n == 0) return;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(sig.arrive, 1u);
        if (prev == gridDim.x - 1) {
            *sig.arrive = 0;                       // the next launch is stream-ordered after this one
            __threadfence_system();
            for (int k = 0; k < sig.n; ++k) st_release_sys_u64(sig.slot[k], sig.value);
        }
    }
}

// MUFU.RSQ, no denormal fix-up code around it
__device__ __forceinline__ float rsqrt_approx(float x)
{
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

} // namespace nb
