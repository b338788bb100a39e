// Micro-probe: throughput of FFMA vs FFMA2 (fma.rn.f32x2) on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters, float w)
{
    float2 a[8];
    for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x + i, threadIdx.x - i);
    float2 ww = make_float2(w, w);
    float2 q = make_float2(1.0001f, 0.9999f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 16; ++r)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) { a[i].x = fmaf(w, q.x, a[i].x); a[i].y = fmaf(w, q.y, a[i].y); }
                else a[i] = __ffma2_rn(ww, q, a[i]);
            }
    }
    float s = 0;
    for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main()
{
    float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 2; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148 * 8, 256>>>(d, 4000, 0.5f); else k<1><<<148 * 8, 256>>>(d, 4000, 0.5f);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double fma = 148.0 * 8 * 256 * 4000.0 * 16 * 8 * 2;
            printf("mode %d (%s): %.3f ms, %.1f GFMA/s, %.1f FMA/clk/SM @1.965GHz\n", mode, mode ? "FFMA2" : "FFMA", ms, fma / ms / 1e6, fma / (ms * 1e-3) / 148 / 1.965e9);
        }
    }
    return 0;
}
