// tc_gemm_probe.cu -- stand-alone check of the tcgen05 building blocks used by kernels_tc.cuh:
// one CTA computes D[128 x N] = A[128 x K] . B[N x K]^T with the 3xTF32 split, operands written to shared memory by
// the threads themselves in the no-swizzle K-major core-matrix layout, result read back with tcgen05.ld.
// Prints the max abs error against a double-precision CPU product for each (N, K) of the model table and for
// both readings of the descriptor's LBO/SBO fields (mode 0 is the one kernels_tc.cuh uses).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc_gemm_probe tools/tc_gemm_probe.cu && ./tc_gemm_probe
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../aidadsp-lv2_b200/csrc/tc_prims.cuh"

using namespace aidax::tc;

__global__ void __launch_bounds__(128) probe_kernel(const float* A, const float* B, float* D, int N, int KP, int swap_fields, int split, int four, int rnd)
{
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int r = threadIdx.x;
    float* Ah = smem;
    float* Al = Ah + 128 * KP;
    float* Bh = Al + 128 * KP;
    float* Bl = Bh + N * KP;
    for (int k = 0; k < KP; ++k) {
        float hi, lo;
        split_tf32(A[r * KP + k], hi, lo);
        if (rnd) lo = __uint_as_float((__float_as_uint(lo) + 0x1000u) & 0xFFFFE000u);
        if (!split) { hi = A[r * KP + k]; lo = 0.f; }
        Ah[(k / 4) * 512 + r * 4 + (k % 4)] = hi;
        Al[(k / 4) * 512 + r * 4 + (k % 4)] = lo;
    }
    for (int idx = r; idx < N * KP; idx += 128) {
        const int n = idx % N, k = idx / N;
        float hi, lo;
        split_tf32(B[n * KP + k], hi, lo);
        if (rnd) lo = __uint_as_float((__float_as_uint(lo) + 0x1000u) & 0xFFFFE000u);
        if (!split) { hi = B[n * KP + k]; lo = 0.f; }
        Bh[(k / 4) * N * 4 + n * 4 + (k % 4)] = hi;
        Bl[(k / 4) * N * 4 + n * 4 + (k % 4)] = lo;
    }
    int cols = 32;
    while (cols < N) cols *= 2;
    if (r == 0) { mbar_init(&bar, 1); mbar_init_fence(); }
    if (r < 32) tmem_alloc(&tmem_slot, cols);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (r == 0) {
        const uint32_t idesc = make_idesc_tf32(128, N);
        const uint32_t lbo_a = 128 * 16, lbo_b = N * 16, sbo = 128;
        uint32_t acc = 0;
        for (int ks = 0; ks < KP / 8; ++ks) {
            const uint32_t ao = ks * 2 * lbo_a, bo = ks * 2 * lbo_b;
            auto desc = [&](const float* base, uint32_t off, uint32_t lbo) {
                return swap_fields ? make_smem_desc(smem_u32(base) + off, sbo, lbo) : make_smem_desc(smem_u32(base) + off, lbo, sbo);
            };
            if (split) {
                if (four) { mma_tf32(tmem, desc(Al, ao, lbo_a), desc(Bl, bo, lbo_b), idesc, acc); acc = 1; }
                mma_tf32(tmem, desc(Al, ao, lbo_a), desc(Bh, bo, lbo_b), idesc, acc); acc = 1;
                mma_tf32(tmem, desc(Ah, ao, lbo_a), desc(Bl, bo, lbo_b), idesc, acc);
            }
            mma_tf32(tmem, desc(Ah, ao, lbo_a), desc(Bh, bo, lbo_b), idesc, acc); acc = 1;
        }
        mma_commit(smem_u32(&bar));
    }
    mbar_wait(smem_u32(&bar), 0);
    tc_fence_after();
    const uint32_t lane_base = tmem + ((uint32_t)((r >> 5) * 32) << 16);
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(lane_base + c0, v);
        tmem_ld_wait16(v);
        for (int i = 0; i < 16; ++i) D[r * N + c0 + i] = __uint_as_float(v[i]);
    }
    tc_fence_before();
    __syncthreads();
    if (r < 32) tmem_dealloc(tmem, cols);
}

int main()
{
    const int Hs[] = {8, 12, 16, 20, 24, 32, 40, 64};
    int bad = 0;
    // pass 0: the layout kernels_tc.cuh uses (modes 0 and 2); pass 1 (diagnostic, only if pass 0 failed): swapped fields
    for (int pass = 0; pass < 2; ++pass) {
    if (pass == 1 && !bad) break;
    for (int H : Hs) {
        const int N = 4 * H, KP = (H + 4 + 7) / 8 * 8;
        std::vector<float> A(128 * KP), B(N * KP), D(128 * N);
        srand(1234 + H);
        for (auto& v : A) v = (float)rand() / RAND_MAX * 2.f - 1.f;
        for (auto& v : B) v = ((float)rand() / RAND_MAX * 2.f - 1.f) * 0.5f;
        float *dA, *dB, *dD;
        cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
        cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
        const size_t smem = (size_t)(2 * 128 * KP + 2 * N * KP) * 4;
        cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        for (int mode = 0; mode < 6; ++mode) {   // 0: lbo=K stride, split ; 1: fields swapped, split ; 2: lbo=K stride, single tf32
            if ((mode == 1) != (pass == 1)) continue;  // 3: four products ; 4: lo rounded to nearest ; 5: both
            cudaMemset(dD, 0xff, D.size() * 4);
            probe_kernel<<<1, 128, smem>>>(dA, dB, dD, N, KP, mode == 1, mode != 2, mode == 3 || mode == 5, mode >= 4);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("H=%d mode=%d CUDA error: %s\n", H, mode, cudaGetErrorString(e)); return 2; }
            cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
            double worst = 0, sq = 0;
            for (int r = 0; r < 128; ++r)
                for (int n = 0; n < N; ++n) {
                    double ref = 0;
                    for (int k = 0; k < KP; ++k) ref += (double)A[r * KP + k] * (double)B[n * KP + k];
                    const double d = fabs(ref - (double)D[r * N + n]);
                    if (!(d <= worst)) worst = d;   // NaN-propagating max
                    sq += d * d;
                }
            const char* names[] = {"lbo=K-stride, 3xTF32", "lbo/sbo swapped, 3xTF32", "lbo=K-stride, 1xTF32", "4 products", "3 products, lo rounded", "4 products, lo rounded"};
            printf("H=%2d N=%3d K=%2d mode=%d (%s) max abs err %.3g rms %.3g\n", H, N, KP, mode, names[mode], worst, sqrt(sq / (128.0 * N)));
            if (mode == 0 && !(worst < 1e-5)) bad = 1;
        }
        cudaFree(dA); cudaFree(dB); cudaFree(dD);
    }
    }
    printf(bad ? "PROBE FAILED\n" : "PROBE OK\n");
    return bad;
}
