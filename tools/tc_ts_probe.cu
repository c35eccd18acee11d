// tc_ts_probe.cu -- stand-alone check of the "TS" form of tcgen05.mma (A operand in TENSOR MEMORY, written by the
// threads with tcgen05.st; B in shared memory) before kernels_tc2.cuh relies on it: one CTA computes
// D[128 x N] = A[128 x K] . B[N x K]^T with the 3xTF32 split and prints the max abs error against a double-precision
// CPU product, for every (N, K) of the model table.  D at TMEM columns [0, N), A_hi at [256, 256+K), A_lo after it.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/tc_ts_probe tools/tc_ts_probe.cu && tools/bin/tc_ts_probe
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../aidadsp-lv2_b200/csrc/tc_prims.cuh"

using namespace aidax::tc;

__global__ void __launch_bounds__(128) probe_kernel(const float* A, const float* B, float* D, float* Aback, int N, int KP)
{
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int r = threadIdx.x;
    float* Bh = smem;
    float* Bl = Bh + N * KP;
    for (int idx = r; idx < N * KP; idx += 128) {
        const int n = idx % N, k = idx / N;
        float hi, lo;
        split_tf32_rn(B[n * KP + k], hi, lo);
        Bh[(k / 4) * N * 4 + n * 4 + (k % 4)] = hi;
        Bl[(k / 4) * N * 4 + n * 4 + (k % 4)] = lo;
    }
    if (r == 0) { mbar_init(&bar, 1); mbar_init_fence(); }
    if (r < 32) tmem_alloc(&tmem_slot, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t lane_base = tmem + ((uint32_t)((r >> 5) * 32) << 16);
    const uint32_t a_hi = 256, a_lo = 256 + KP;
    // every thread writes its own row of A (hi, lo) into tensor memory, 4 columns per store
    for (int k = 0; k < KP; k += 4) {
        float h[4], l[4];
        for (int i = 0; i < 4; ++i) split_tf32(A[r * KP + k + i], h[i], l[i]);
        tmem_st4(lane_base + a_hi + k, __float_as_uint(h[0]), __float_as_uint(h[1]), __float_as_uint(h[2]), __float_as_uint(h[3]));
        tmem_st4(lane_base + a_lo + k, __float_as_uint(l[0]), __float_as_uint(l[1]), __float_as_uint(l[2]), __float_as_uint(l[3]));
    }
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    if (r == 0) {
        tc_fence_after();
        const uint32_t idesc = make_idesc_tf32(128, N);
        const uint32_t lbo_b = N * 16, sbo = 128;
        uint32_t acc = 0;
        for (int ks = 0; ks < KP / 8; ++ks) {
            const uint32_t bo = ks * 2 * lbo_b;
            mma_tf32_ts(tmem, tmem + a_lo + ks * 8, make_smem_desc(smem_u32(Bh) + bo, lbo_b, sbo), idesc, acc); acc = 1;
            mma_tf32_ts(tmem, tmem + a_hi + ks * 8, make_smem_desc(smem_u32(Bl) + bo, lbo_b, sbo), idesc, acc);
            mma_tf32_ts(tmem, tmem + a_hi + ks * 8, make_smem_desc(smem_u32(Bh) + bo, lbo_b, sbo), idesc, acc);
        }
        mma_commit(smem_u32(&bar));
    }
    mbar_wait(smem_u32(&bar), 0);
    tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(lane_base + c0, v);
        tmem_ld_wait16(v);
        for (int i = 0; i < 16; ++i) D[r * N + c0 + i] = __uint_as_float(v[i]);
    }
    for (int k = 0; k < KP; k += 4) { // read the operand back: what store_state of the kernel will do
        uint32_t h[4], l[4];
        tmem_ld4(lane_base + a_hi + k, h);
        tmem_ld4(lane_base + a_lo + k, l);
        for (int i = 0; i < 4; ++i) Aback[r * KP + k + i] = __uint_as_float(h[i]) + __uint_as_float(l[i] & 0xFFFFE000u);
    }
    tc_fence_before();
    __syncthreads();
    if (r < 32) tmem_dealloc(tmem, 512);
}

int main()
{
    const int Hs[] = {8, 12, 16, 20, 24, 32, 40};
    int bad = 0;
    for (int H : Hs) {
        const int N = 4 * H, KP = (H + 4 + 7) / 8 * 8;
        std::vector<float> A(128 * KP), B(N * KP), D(128 * N), Ab(128 * KP);
        srand(1234 + H);
        for (auto& v : A) v = (float)rand() / RAND_MAX * 2.f - 1.f;
        for (auto& v : B) v = ((float)rand() / RAND_MAX * 2.f - 1.f) * 0.5f;
        float *dA, *dB, *dD, *dAb;
        cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4); cudaMalloc(&dAb, Ab.size() * 4);
        cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
        cudaMemset(dD, 0xff, D.size() * 4);
        const size_t smem = (size_t)(2 * N * KP) * 4;
        cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        probe_kernel<<<1, 128, smem>>>(dA, dB, dD, dAb, N, KP);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("H=%d CUDA error: %s\n", H, cudaGetErrorString(e)); return 2; }
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(Ab.data(), dAb, Ab.size() * 4, cudaMemcpyDeviceToHost);
        double worst = 0, sq = 0, wback = 0;
        for (int r = 0; r < 128; ++r)
            for (int n = 0; n < N; ++n) {
                double ref = 0;
                for (int k = 0; k < KP; ++k) ref += (double)A[r * KP + k] * (double)B[n * KP + k];
                const double d = fabs(ref - (double)D[r * N + n]);
                if (!(d <= worst)) worst = d;
                sq += d * d;
            }
        for (size_t i = 0; i < A.size(); ++i) { const double d = fabs((double)A[i] - (double)Ab[i]); if (!(d <= wback)) wback = d; }
        printf("TS form H=%2d N=%3d K=%2d max abs err %.3g rms %.3g | operand read-back max err %.3g\n", H, N, KP, worst, sqrt(sq / (128.0 * N)), wback);
        if (!(worst < 1e-5) || !(wback < 1e-6)) bad = 1;
        cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dAb);
    }
    printf(bad ? "TS PROBE FAILED\n" : "TS PROBE OK\n");
    return bad;
}
