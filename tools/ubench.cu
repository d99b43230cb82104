// tools/ubench.cu -- pipe-rate microbenchmarks behind the kernel design (NOT product code).
// Measures warp-instructions per cycle per SM sub-partition for packed fp32 (FFMA2) with different
// operand patterns, alone and mixed with MUFU.RSQ / LDS, on sm_100a.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ float rsq(float x) { float y; asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

constexpr int ITERS = 4096;
constexpr int K = 8; // independent chains per thread

// MODE 0: FFMA2 3 distinct pairs      acc[k] = a[k]*b[k]+acc[k]
// MODE 1: FFMA2 shared b (reuse)      acc[k] = a[k]*b   +acc[k]
// MODE 2: FFMA2 2 distinct pairs      acc[k] = a[k]*a[k]+acc[k]
// MODE 3: scalar FFMA 3 distinct      acc[k] = a[k]*b[k]+acc[k]   (2 per chain to match flops)
// MODE 4: MODE 1 + 1 MUFU per 6 FFMA2 (MUFU result feeds nothing on the critical path)
// MODE 5: MODE 0 + 1 MUFU per 6 FFMA2
// MODE 6: MUFU only
// MODE 7: MODE 1 + 1 MUFU per 3 FFMA2
// MODE 9 / 10: MODE 2 (FFMA2 a*a+acc, 2 cycles each) + 12 / 24 scalar FFMA per iteration: do packed and
//               scalar fp32 share one datapath, or can they overlap?
// MODE 8: FFMA2 scalar-broadcast c     acc[k] = a[k]*a[k] + s  then acc used next iter as a
template <int MODE>
__global__ void __launch_bounds__(256) ub(float2 *out, float seed, long long *cycles)
{
    float2 a[K], b[K], acc[K];
    float m[K];
    float sc[K], sa[K];
    for (int k = 0; k < K; ++k) { sc[k] = seed * k; sa[k] = 1.0f + 1e-5f * k + seed * 1e-6f; }
    for (int k = 0; k < K; ++k) {
        a[k] = make_float2(seed + k + threadIdx.x * 1e-3f, seed - k);
        b[k] = make_float2(1.0f + 1e-6f * k, 1.0f - 1e-6f * k);
        acc[k] = make_float2(0.f, 0.f);
        m[k] = 1.0f + k + seed;
    }
    const float2 bs = make_float2(1.0f + seed * 1e-7f, 1.0f - seed * 1e-7f);
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int r = 0; r < 3; ++r) {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                if (MODE == 0 || MODE == 5) acc[k] = __ffma2_rn(a[k], b[k], acc[k]);
                if (MODE == 1 || MODE == 4 || MODE == 7) acc[k] = __ffma2_rn(a[k], bs, acc[k]);
                if (MODE == 2) acc[k] = __ffma2_rn(a[k], a[k], acc[k]);
                if (MODE == 3) { acc[k].x = fmaf(a[k].x, b[k].x, acc[k].x); acc[k].y = fmaf(a[k].y, b[k].y, acc[k].y); }
                if (MODE == 8) acc[k] = __ffma2_rn(acc[k], a[k], make_float2(seed, seed));
                if (MODE == 9 || MODE == 10) acc[k] = __ffma2_rn(a[k], a[k], acc[k]);
                if (MODE == 10 || (MODE == 9 && (k & 1))) sc[k] = fmaf(sc[k], sa[k], sa[k]);
            }
            if (MODE == 4 || MODE == 5) {
#pragma unroll
                for (int k = 0; k < K / 6 + 1; ++k) if (r * (K / 6 + 1) + k < 4) m[r * 2 + k] = rsq(m[r * 2 + k]);
            }
            if (MODE == 7) {
#pragma unroll
                for (int k = 0; k < 3; ++k) m[(r * 3 + k) % K] = rsq(m[(r * 3 + k) % K]);
            }
        }
        if (MODE == 6) {
#pragma unroll
            for (int k = 0; k < K; ++k) m[k] = rsq(m[k]);
        }
    }
    long long t1 = clock64();
    float2 s = make_float2(0.f, 0.f);
    for (int k = 0; k < K; ++k) { s.x += acc[k].x + m[k]; s.y += acc[k].y + sc[k]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int MODE>
void run(const char *name, int ctas_per_sm, double fma_per_iter, double mufu_per_iter)
{
    float2 *out; long long *cyc, h;
    const int grid = 148 * ctas_per_sm;
    CK(cudaMalloc(&out, (size_t)grid * 256 * sizeof(float2)));
    CK(cudaMalloc(&cyc, 8));
    ub<MODE><<<grid, 256>>>(out, 1.0f, cyc);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    ub<MODE><<<grid, 256>>>(out, 1.0f, cyc);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    CK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
    // per SMSP: warps = ctas_per_sm * 8 / 4
    const double warps_per_smsp = ctas_per_sm * 2.0;
    const double cyc_per_iter = (double)h / ITERS;                 // one CTA's view
    const double fma_rate = fma_per_iter * warps_per_smsp / cyc_per_iter;   // warp-inst / clk / SMSP
    const double mufu_rate = mufu_per_iter * warps_per_smsp / cyc_per_iter;
    printf("%-40s ctas/sm=%d  %8.1f cyc/iter  FMA-pipe inst/clk/SMSP=%.3f (x2 lanes-cycles: %.1f%%)  MUFU/clk/SMSP=%.4f (%.1f%% of 1/8)  %.3f ms\n",
           name, ctas_per_sm, cyc_per_iter, fma_rate, MODE == 3 ? fma_rate * 100 : fma_rate * 200, mufu_rate, mufu_rate * 800, ms);
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    for (int c : {1}) {
        run<0>("FFMA2 3 distinct pairs", c, 24, 0);
        run<1>("FFMA2 shared b (reuse)", c, 24, 0);
        run<2>("FFMA2 a*a+acc (2 pairs)", c, 24, 0);
        run<8>("FFMA2 acc*a+scalar", c, 24, 0);
        run<3>("FFMA scalar 3 distinct (48/iter)", c, 48, 0);
        run<4>("FFMA2 shared b + MUFU 1:6", c, 24, 4);
        run<5>("FFMA2 3 distinct + MUFU 1:6", c, 24, 4);
        run<7>("FFMA2 shared b + MUFU 3:8", c, 24, 9);
        run<6>("MUFU only", c, 0, 8);
        run<9>("FFMA2 a*a+acc + 12 scalar FFMA (x2 col = packed only)", c, 24, 0);
        run<10>("FFMA2 a*a+acc + 24 scalar FFMA (x2 col = packed only)", c, 24, 0);
    }
    return 0;
}
