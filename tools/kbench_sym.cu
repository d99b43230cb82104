// tools/kbench_sym.cu -- EXPERIMENT (not product code): all-pairs with Newton's third law.
// Each unordered pair of bodies is evaluated once; the contribution goes to the target (register
// accumulators, as in the product kernel) AND, negated, to the source (per-lane partial sums, reduced
// across the warp with shuffles, accumulated in shared memory per stage, flushed with atomics).
// Question: does halving the pair evaluations beat the cost of the reduction on B200?
#include "../nbodysim_b200/csrc/force_f32_fast.cuh"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
using namespace nb;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int SY_THREADS = 256, SY_STAGE = 2; // stage = 2 blocks = 512 sources

// reduce V packed values over the warp with the transposing butterfly: after the call, lane l holds the
// total of value (l >> shift...) -- here simply a full butterfly for clarity (V * 5 shuffles)
__device__ __forceinline__ float2 warp_sum2(float2 v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
        v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
    }
    return v;
}

template <int I, bool SYM>
__global__ void __launch_bounds__(SY_THREADS, 1)
sym_kernel(const float *__restrict__ posm, float *__restrict__ acc, int ntiles, float eps2)
{
    constexpr int TILE_BLKS = I;                           // 256 threads: one block per target slot
    using RingT = Ring<BLK_ELEMS, SY_STAGE>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    RingT ring;
    ring.setup(smem_raw, SY_THREADS / 32);
    float *sacc = reinterpret_cast<float *>(smem_raw + RingT::SMEM);   // [3][512] source-side partial sums

    // tile pair (A <= B) from the linear block index (row-major over the upper triangle incl. diagonal)
    int A, B;
    if (SYM) {
        long long p = blockIdx.x, row = 0, rem = ntiles;
        // rows have ntiles, ntiles-1, ... entries
        row = (long long)floor(((2.0 * ntiles + 1) - sqrt((2.0 * ntiles + 1) * (2.0 * ntiles + 1) - 8.0 * (double)p)) / 2.0);
        while (row * ntiles - row * (row - 1) / 2 > p) --row;
        while ((row + 1) * ntiles - (row + 1) * row / 2 <= p) ++row;
        rem = p - (row * ntiles - row * (row - 1) / 2);
        A = (int)row; B = (int)(row + rem);
    } else {
        A = blockIdx.x / ntiles; B = blockIdx.x % ntiles;
    }
    const bool offdiag = SYM && (A != B);
    const int tid = threadIdx.x, lane = tid & 31;
    const int jb0 = B * TILE_BLKS, chunk_blks = TILE_BLKS;
    const int nst = (chunk_blks + SY_STAGE - 1) / SY_STAGE;
    const float *src = posm + (size_t)jb0 * BLK_ELEMS;
    if (tid == 0) for (int t = 0; t < min(NSTAGE, nst); ++t) ring.issue(src, t, chunk_blks);

    float2 nxi[I], nyi[I], nzi[I], ax[I], ay[I], az[I];
#pragma unroll
    for (int k = 0; k < I; ++k) {
        const float *b = posm + (size_t)(A * TILE_BLKS + k) * BLK_ELEMS + tid;
        const float x = b[0], y = b[BLK], z = b[2 * BLK];
        nxi[k] = make_float2(-x, -x); nyi[k] = make_float2(-y, -y); nzi[k] = make_float2(-z, -z);
        ax[k] = ay[k] = az[k] = make_float2(0.f, 0.f);
    }
    const float2 e2 = make_float2(eps2, eps2);
    for (int q = tid; q < 3 * SY_STAGE * BLK; q += SY_THREADS) sacc[q] = 0.f;
    __syncthreads();

    for (int t = 0; t < nst; ++t) {
        const int s = t % NSTAGE;
        mbar_wait(&ring.full[s], (uint32_t)(t / NSTAGE) & 1u);
        const float *st = ring.stage + (size_t)s * RingT::STAGE_FLOATS;
        const int nb = min(SY_STAGE, chunk_blks - t * SY_STAGE);
        for (int b = 0; b < nb; ++b) {
            const float *sx = st + b * BLK_ELEMS;
#pragma unroll 1
            for (int j = 0; j < BLK; j += 2) {
                const float2 xj = *reinterpret_cast<const float2 *>(sx + j);
                const float2 yj = *reinterpret_cast<const float2 *>(sx + BLK + j);
                const float2 zj = *reinterpret_cast<const float2 *>(sx + 2 * BLK + j);
                float2 px = make_float2(0.f, 0.f), py = px, pz = px;      // source-side partial sums of this lane
#pragma unroll
                for (int k = 0; k < I; ++k) {
                    const float2 dx = __fadd2_rn(xj, nxi[k]);
                    const float2 dy = __fadd2_rn(yj, nyi[k]);
                    const float2 dz = __fadd2_rn(zj, nzi[k]);
                    float2 r2 = __ffma2_rn(dx, dx, e2);
                    r2 = __ffma2_rn(dy, dy, r2);
                    r2 = __ffma2_rn(dz, dz, r2);
                    const float2 ri = make_float2(rsqrt_approx(r2.x), rsqrt_approx(r2.y));
                    const float2 sc = __fmul2_rn(__fmul2_rn(ri, ri), ri);
                    ax[k] = __ffma2_rn(dx, sc, ax[k]);
                    ay[k] = __ffma2_rn(dy, sc, ay[k]);
                    az[k] = __ffma2_rn(dz, sc, az[k]);
                    if (SYM) {
                        px = __ffma2_rn(dx, sc, px);
                        py = __ffma2_rn(dy, sc, py);
                        pz = __ffma2_rn(dz, sc, pz);
                    }
                }
                if (offdiag) {   // warp-uniform
                    px = warp_sum2(px); py = warp_sum2(py); pz = warp_sum2(pz);
                    if (lane == 0) {
                        float *o = sacc + b * BLK + j;
                        atomicAdd(o, px.x); atomicAdd(o + 1, px.y);
                        atomicAdd(o + SY_STAGE * BLK, py.x); atomicAdd(o + SY_STAGE * BLK + 1, py.y);
                        atomicAdd(o + 2 * SY_STAGE * BLK, pz.x); atomicAdd(o + 2 * SY_STAGE * BLK + 1, pz.y);
                    }
                }
            }
        }
        if (offdiag) {
            __syncthreads();   // all warps have added their partial sums of this stage
            for (int q = tid; q < 3 * nb * BLK; q += SY_THREADS) {
                const int c = q / (nb * BLK), r = q % (nb * BLK);
                const int blk = jb0 + t * SY_STAGE + r / BLK, l = r % BLK;
                atomicAdd(acc + (size_t)blk * BLK_ELEMS + c * BLK + l, -sacc[c * SY_STAGE * BLK + r]);   // a_j -= sum
                sacc[c * SY_STAGE * BLK + r] = 0.f;
            }
            __syncthreads();
        }
        ring.release_and_refill(src, t, nst, chunk_blks);
    }
#pragma unroll
    for (int k = 0; k < I; ++k) {
        float *o = acc + (size_t)(A * TILE_BLKS + k) * BLK_ELEMS + tid;
        atomicAdd(o, ax[k].x + ax[k].y); atomicAdd(o + BLK, ay[k].x + ay[k].y); atomicAdd(o + 2 * BLK, az[k].x + az[k].y);
    }
}

template <int I, bool SYM>
float run(const char *name, const float *posm, float *acc, int nblk, float eps2, int reps, std::vector<float> &out)
{
    auto kern = sym_kernel<I, SYM>;
    using RingT = Ring<BLK_ELEMS, SY_STAGE>;
    const size_t smem = RingT::SMEM + 3 * SY_STAGE * BLK * sizeof(float);
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int ntiles = nblk / I;
    const long long grid = SYM ? (long long)ntiles * (ntiles + 1) / 2 : (long long)ntiles * ntiles;
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, kern));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaMemset(acc, 0, (size_t)nblk * BLK_ELEMS * 4));
    kern<<<(unsigned)grid, SY_THREADS, smem>>>(posm, acc, ntiles, eps2);
    CK(cudaDeviceSynchronize());
    out.resize((size_t)nblk * BLK_ELEMS);
    CK(cudaMemcpy(out.data(), acc, out.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; ++r) kern<<<(unsigned)grid, SY_THREADS, smem>>>(posm, acc, ntiles, eps2);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= reps;
    const double n = (double)nblk * BLK;
    printf("%-34s regs=%3d grid=%8lld  %9.3f ms  %8.1f G inter/s  (%5.1f%% of 74.45 TF at 20 flop)\n", name, fa.numRegs, grid, ms,
           n * n / (ms * 1e-3) / 1e9, 100.0 * n * n / (ms * 1e-3) * 20 / 74.45e12);
    fflush(stdout);
    return ms;
}

int main(int argc, char **argv)
{
    int n = argc > 1 ? atoi(argv[1]) : 262144;
    int reps = argc > 2 ? atoi(argv[2]) : 3;
    n = (n / 6144) * 6144;
    const int nblk = n / BLK;
    std::vector<float> h((size_t)nblk * BLK_ELEMS);
    srand(1);
    auto U = []() { return (float)rand() / (float)RAND_MAX; };
    for (int i = 0; i < n; ++i) {
        float x, y, z;
        do { x = 2 * U() - 1; y = 2 * U() - 1; z = 2 * U() - 1; } while (x * x + y * y + z * z > 1);
        size_t o = blk_index(i, 0);
        h[o] = x; h[o + BLK] = y; h[o + 2 * BLK] = z; h[o + 3 * BLK] = 1.0f;
    }
    float *posm, *acc;
    CK(cudaMalloc(&posm, h.size() * 4)); CK(cudaMalloc(&acc, h.size() * 4));
    CK(cudaMemcpy(posm, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    printf("N=%d\n", n);
    std::vector<float> ref, out;
    run<8, false>("direct    I8 (atomics epilogue)", posm, acc, nblk, 1e-4f, reps, ref);
    run<8, true>("symmetric I8", posm, acc, nblk, 1e-4f, reps, out);
    double maxrel = 0;
    for (int b = 0; b < nblk; ++b) for (int l = 0; l < BLK; ++l) {
        size_t o = (size_t)b * BLK_ELEMS + l;
        double dx = out[o] - ref[o], dy = out[o + BLK] - ref[o + BLK], dz = out[o + 2 * BLK] - ref[o + 2 * BLK];
        double nn = sqrt((double)ref[o] * ref[o] + (double)ref[o + BLK] * ref[o + BLK] + (double)ref[o + 2 * BLK] * ref[o + 2 * BLK]);
        maxrel = fmax(maxrel, sqrt(dx * dx + dy * dy + dz * dz) / (nn + 1e-30));
    }
    printf("symmetric vs direct: max relative difference %.2e\n", maxrel);
    run<6, true>("symmetric I6", posm, acc, nblk, 1e-4f, reps, out);
    run<4, true>("symmetric I4", posm, acc, nblk, 1e-4f, reps, out);
    return 0;
}
