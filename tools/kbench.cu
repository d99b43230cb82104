// tools/kbench.cu -- tuning harness for the all-pairs force kernel (NOT part of the product path).
// Times template variants of force_f32_fast_kernel on synthetic sources with CUDA events and checks
// every variant against variant 0.  Build: make kbench ; run on the GPU box: build/kbench [N] [reps]
#include "../nbodysim_b200/csrc/force_f32_fast.cuh"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace nb;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

struct Result { float ms; double maxrel; };

template <int I, int THREADS, int MINB, int UNROLL, int STAGE_BLKS, int FORM, int MINCHUNK = 8>
Result run(const char *name, const float *posm, float *accp, int nblk, int splits, float eps2,
           int reps, const std::vector<float> &ref, std::vector<float> &out, int sms)
{
    auto kern = force_f32_fast_kernel<I, THREADS, MINB, UNROLL, STAGE_BLKS, FORM, false, false>;
    using RingT = Ring<BLK_ELEMS, STAGE_BLKS>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RingT::SMEM));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, RingT::SMEM));
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, kern));
    constexpr int TILE_BLKS = I / (BLK / THREADS);
    FastArgs a{};
    a.posm = posm; a.accp = accp;
    a.i_blk0 = 0; a.i_blk_local0 = 0; a.n_iblk_shard = nblk; a.j_blk0 = 0; a.j_nblk = nblk;
    if (splits <= 0) { // pick splits for ~whole waves
        int tiles = nblk / TILE_BLKS, slots = sms * occ; double best = -1; splits = 1;
        for (int s = 1; s <= 64 && nblk / s >= MINCHUNK; ++s) {
            long long units = (long long)tiles * s, waves = (units + slots - 1) / slots;
            double eff = (double)units / (waves * slots); if (waves < 4) eff *= 0.97;
            if (eff > best * 1.01) { best = eff; splits = s; }
        }
    }
    a.splits = splits; a.slot0 = 0; a.eps2 = eps2; a.acc_scale = 1.0f; a.n_real = (long long)nblk * BLK;
    const int grid = (nblk / TILE_BLKS) * splits;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int w = 0; w < 2; ++w) kern<<<grid, THREADS, RingT::SMEM>>>(a);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; ++r) kern<<<grid, THREADS, RingT::SMEM>>>(a);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= reps;
    // reduce partial slots on host for the check
    size_t nb_e = (size_t)nblk * BLK_ELEMS;
    std::vector<float> h(nb_e * splits);
    CK(cudaMemcpy(h.data(), accp, h.size() * 4, cudaMemcpyDeviceToHost));
    out.assign(nb_e, 0.f);
    for (int s = 0; s < splits; ++s) for (size_t k = 0; k < nb_e; ++k) out[k] += h[s * nb_e + k];
    double maxrel = 0;
    if (!ref.empty()) {
        for (int b = 0; b < nblk; ++b) for (int l = 0; l < BLK; ++l) {
            size_t o = (size_t)b * BLK_ELEMS + l;
            double dx = out[o] - ref[o], dy = out[o + BLK] - ref[o + BLK], dz = out[o + 2 * BLK] - ref[o + 2 * BLK];
            double nn = sqrt((double)ref[o] * ref[o] + (double)ref[o + BLK] * ref[o + BLK] + (double)ref[o + 2 * BLK] * ref[o + 2 * BLK]);
            double r = sqrt(dx * dx + dy * dy + dz * dz) / (nn + 1e-30);
            if (r > maxrel) maxrel = r;
        }
    }
    const double n = (double)nblk * BLK;
    const double ginter = n * n / (ms * 1e-3) / 1e9;
    printf("%-44s regs=%3d occ=%d splits=%2d grid=%6d  %8.3f ms  %8.1f G/s  %5.1f%% of 74.45TF  maxrel=%.2e\n", name,
           fa.numRegs, occ, splits, grid, ms, ginter, 100.0 * ginter * 20.0 / 74449.92, maxrel);
    fflush(stdout);
    return {ms, maxrel};
}

int main(int argc, char **argv)
{
    int n = argc > 1 ? atoi(argv[1]) : 262144;
    int reps = argc > 2 ? atoi(argv[2]) : 5;
    n = (n / 6144) * 6144;   // whole tiles for every I in {1,2,4,6,8}
    const int nblk = n / BLK;
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    printf("device %s, %d SMs, N=%d, reps=%d\n", prop.name, prop.multiProcessorCount, n, reps);
    std::vector<float> h((size_t)nblk * BLK_ELEMS);
    srand(1);
    auto U = []() { return (float)rand() / (float)RAND_MAX; };
    for (int i = 0; i < n; ++i) {
        float x, y, z;
        do { x = 2 * U() - 1; y = 2 * U() - 1; z = 2 * U() - 1; } while (x * x + y * y + z * z > 1);
        size_t o = blk_index(i, 0);
        h[o] = x; h[o + BLK] = y; h[o + 2 * BLK] = z; h[o + 3 * BLK] = 1.0f;  // equal masses (so the uniform-mass form can be checked too)
    }
    float *posm, *accp;
    CK(cudaMalloc(&posm, h.size() * 4));
    CK(cudaMalloc(&accp, h.size() * 4 * 64));
    CK(cudaMemcpy(posm, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    const float eps2 = 1e-4f;
    CK(cudaDeviceSynchronize());
    const int sms = prop.multiProcessorCount;
    std::vector<float> ref, out, none;
    //            I  THR MINB UNR STG FORM (0 = general masses, 1 = uniform mass) [MINCHUNK]
    if (n >= 60000) {
        run<8, 256, 1, 1, 2, 0>("plain   I8 t256 b1 u1 s2 (product)", posm, accp, nblk, 0, eps2, reps, none, ref, sms);
        run<8, 256, 1, 1, 2, 1>("uniform I8 t256 b1 u1 s2 (product)", posm, accp, nblk, 0, eps2, reps, ref, out, sms);
        run<4, 256, 2, 2, 2, 0>("plain   I4 t256 b2 u2 s2", posm, accp, nblk, 0, eps2, reps, ref, out, sms);
        run<4, 256, 2, 1, 2, 1>("uniform I4 t256 b2 u1 s2", posm, accp, nblk, 0, eps2, reps, ref, out, sms);
        run<8, 128, 2, 1, 2, 1>("uniform I8 t128 b2 u1 s2", posm, accp, nblk, 0, eps2, reps, ref, out, sms);
        run<6, 128, 3, 1, 2, 1>("uniform I6 t128 b3 u1 s2", posm, accp, nblk, 0, eps2, reps, ref, out, sms);
    } else { // small shards: small-tile geometries, source chunks down to one block
        run<2, 128, 6, 2, 1, 1, 1>("uniform I2 t128 b6 u2 s1 (product small)", posm, accp, nblk, 0, eps2, reps, none, ref, sms);
        run<2, 128, 6, 1, 1, 1, 1>("uniform I2 t128 b6 u1 s1", posm, accp, nblk, 0, eps2, reps, ref, out, sms);
        run<2, 128, 8, 2, 1, 1, 1>("uniform I2 t128 b8 u2 s1", posm, accp, nblk, 0, eps2, reps, ref, out, sms);
        run<4, 128, 4, 1, 1, 1, 1>("uniform I4 t128 b4 u1 s1", posm, accp, nblk, 0, eps2, reps, ref, out, sms);
        run<4, 128, 4, 2, 1, 1, 1>("uniform I4 t128 b4 u2 s1", posm, accp, nblk, 0, eps2, reps, ref, out, sms);
        run<4, 128, 3, 1, 2, 1, 2>("uniform I4 t128 b3 u1 s2", posm, accp, nblk, 0, eps2, reps, ref, out, sms);
        run<4, 256, 2, 1, 1, 1, 1>("uniform I4 t256 b2 u1 s1", posm, accp, nblk, 0, eps2, reps, ref, out, sms);
        run<2, 256, 4, 2, 1, 1, 1>("uniform I2 t256 b4 u2 s1", posm, accp, nblk, 0, eps2, reps, ref, out, sms);
        run<1, 256, 4, 2, 1, 1, 1>("uniform I1 t256 b4 u2 s1", posm, accp, nblk, 0, eps2, reps, ref, out, sms);
        run<6, 128, 3, 1, 1, 1, 1>("uniform I6 t128 b3 u1 s1", posm, accp, nblk, 0, eps2, reps, ref, out, sms);
        run<8, 256, 1, 1, 2, 1, 2>("uniform I8 t256 b1 u1 s2 (large tile)", posm, accp, nblk, 0, eps2, reps, ref, out, sms);
    }
    return 0;
}
