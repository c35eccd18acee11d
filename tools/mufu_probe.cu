// Scratch: measured error of the XU-pipe approximations the gate epilogue is built from, on the GPU itself:
// ex2.approx.ftz.f32 and rcp.approx.ftz.f32 (relative, in fp32 ulps) and the composite sigmoid / tanh forms
// (absolute), against double precision.  Decides where the 10-second error budget of a sensitive model goes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/mufu_probe tools/mufu_probe.cu && tools/bin/mufu_probe
#include <cstdio>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2a(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcpa(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

__device__ float ex2_poly6(float e)
{
    const float ec = fminf(fmaxf(e, -125.0f), 24.0f);
    const float t = ec + 12582912.0f;
    const float f = ec - (t - 12582912.0f);
    float p = 1.534581097e-04f;
    p = fmaf(p, f, 1.339993090e-03f); p = fmaf(p, f, 9.618489072e-03f); p = fmaf(p, f, 5.550328642e-02f);
    p = fmaf(p, f, 2.402264625e-01f); p = fmaf(p, f, 6.931471825e-01f); p = fmaf(p, f, 1.0f);
    return __uint_as_float(__float_as_uint(p) + (__float_as_uint(t) << 23));
}

// mode 0: ex2 (rel, vs exp2(double(x))), 1: rcp (rel), 2: poly6 (rel), 3: sigmoid(x) = rcp(1+ex2(-x*l2e)) (abs),
// 4: tanh(x) = 1 - 2*rcp(1+ex2(2x*l2e)) (abs), 5: sigmoid with one Newton step on rcp, 6: sigmoid with poly ex2 + Newton
__global__ void probe(int mode, float lo, float hi, int n, double* sum2, double* maxe, double* sum1)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double err = 0.0;
    if (i < n) {
        const float x = lo + (hi - lo) * ((float)i + 0.5f) / (float)n;
        const float L2E = 1.4426950408889634f;
        if (mode == 0) { const double r = exp2((double)x); err = ((double)ex2a(x) - r) / r * 8388608.0; }
        else if (mode == 1) { const double r = 1.0 / (double)x; err = ((double)rcpa(x) - r) / r * 8388608.0; }
        else if (mode == 2) { const double r = exp2((double)x); err = ((double)ex2_poly6(x) - r) / r * 8388608.0; }
        else if (mode == 3) { const float a = x * -L2E; err = (double)rcpa(1.0f + ex2a(a)) - 1.0 / (1.0 + exp2((double)a)); }
        else if (mode == 4) { const float a = x * (2.f * L2E); err = (double)fmaf(-2.0f, rcpa(1.0f + ex2a(a)), 1.0f) - (1.0 - 2.0 / (1.0 + exp2((double)a))); }
        else if (mode == 5) { const float a = x * -L2E; const float d = 1.0f + ex2a(a); float r = rcpa(d); r = r * fmaf(-d, r, 2.0f);
                              err = (double)r - 1.0 / (1.0 + exp2((double)a)); }
        else if (mode == 6) { const float a = x * -L2E; const float d = 1.0f + ex2_poly6(a); float r = rcpa(d); r = r * fmaf(-d, r, 2.0f);
                              err = (double)r - 1.0 / (1.0 + exp2((double)a)); }
        else if (mode == 7) { const float a = x * -L2E; const float d = 1.0f + ex2_poly6(a); err = (double)rcpa(d) - 1.0 / (1.0 + exp2((double)a)); }
    }
    // block reduction (simple)
    __shared__ double s2[256], sm[256], s1[256];
    s2[threadIdx.x] = err * err; sm[threadIdx.x] = fabs(err); s1[threadIdx.x] = err;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) { s2[threadIdx.x] += s2[threadIdx.x + s]; sm[threadIdx.x] = fmax(sm[threadIdx.x], sm[threadIdx.x + s]); s1[threadIdx.x] += s1[threadIdx.x + s]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { sum2[blockIdx.x] = s2[0]; maxe[blockIdx.x] = sm[0]; sum1[blockIdx.x] = s1[0]; }
}

int main()
{
    const int n = 1 << 24, nb = n / 256;
    double *d2, *dm, *d1;
    cudaMalloc(&d2, nb * 8); cudaMalloc(&dm, nb * 8); cudaMalloc(&d1, nb * 8);
    std::vector<double> h2(nb), hm(nb), h1(nb);
    const char* names[] = {"ex2.approx rel [ulp]", "rcp.approx rel [ulp]", "poly6 ex2 rel [ulp]", "sigmoid mufu abs", "tanh mufu abs",
                           "sigmoid mufu+newton abs", "sigmoid poly6+newton abs", "sigmoid poly6+rcp abs"};
    for (int mode = 0; mode < 8; ++mode) {
        float lo = -8.f, hi = 8.f;
        if (mode == 1) { lo = 1.f; hi = 4.f; }
        if (mode == 0 || mode == 2) { lo = -16.f; hi = 16.f; }
        probe<<<nb, 256>>>(mode, lo, hi, n, d2, dm, d1);
        cudaMemcpy(h2.data(), d2, nb * 8, cudaMemcpyDeviceToHost); cudaMemcpy(hm.data(), dm, nb * 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(h1.data(), d1, nb * 8, cudaMemcpyDeviceToHost);
        double s2 = 0, m = 0, s1 = 0;
        for (int i = 0; i < nb; ++i) { s2 += h2[i]; m = fmax(m, hm[i]); s1 += h1[i]; }
        printf("%-28s x in [%g, %g]: max %.3e  rms %.3e  mean %.3e\n", names[mode], lo, hi, m, sqrt(s2 / n), s1 / n);
    }
    return cudaGetLastError() != cudaSuccess;
}
