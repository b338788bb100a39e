// tc_stack.cu -- cross-node fusion of the hot chain of a multi-track timeline:
//   per track:  [toucan:Blur] -> [toucan:Dissolve with the neighbouring clip] -> CompNode(premult + over)
// evaluated for ALL tracks in one pass over the output tile, accumulator in registers.
// Compulsory HBM traffic = every referenced source frame read once (+ halo) and the
// final frame written once, instead of one read+write of a full frame per graph node
// (lib/toucanRender/ImageGraph.cpp:165-198, Comp.cpp:29-84, FilterPlugin.cpp:188-202,
// TransitionPlugin.cpp:242-264).  float32 RGBA only (the working format of the 8K config).
#include <cmath>
#include <cstring>
#include <vector>

#include "tc_common.cuh"

namespace tc {

constexpr int ST_MAX_R = 12;     // live taps per side (radius parameter up to ~26)
constexpr int ST_MAX_KERNELS = 4; // distinct blur kernels per launch
constexpr int ST_MAX_LAYERS = 32;
constexpr int ST_TW = 32, ST_TH = 32, ST_THREADS = 256;
constexpr int ST_PER_THREAD = ST_TW * ST_TH / ST_THREADS;

struct StKernel {
    int R;
    float scale;
    float g[2 * ST_MAX_R + 1];
};
struct StLayer {
    const float4* a;
    const float4* b;
    int ka, kb; // kernel index or -1
    float v, iv;
};
struct StParams {
    int w, h, n_layers;
    float4 background;
    StKernel k[ST_MAX_KERNELS];
    StLayer layer[ST_MAX_LAYERS];
};

__device__ __forceinline__ float4 fma4s(float w, float4 p, float4 a)
{
    return make_float4(fmaf(w, p.x, a.x), fmaf(w, p.y, a.y), fmaf(w, p.z, a.z), fmaf(w, p.w, a.w));
}

// Blur one source over the CTA's tile; results for this thread's pixels in out[].
__device__ __forceinline__ void tile_blur(const float4* __restrict__ src, int w, int h, int x0, int y0, const StKernel& K,
                                          float4* in, float4* mid, float4 out[ST_PER_THREAD])
{
    const int R = K.R;
    const int IW = ST_TW + 2 * R, IH = ST_TH + 2 * R;
    __syncthreads(); // previous users of in/mid are done
    for (int i = threadIdx.x; i < IH * IW; i += ST_THREADS) {
        const int iy = i / IW, ix = i - iy * IW;
        const int sx = min(max(x0 - R + ix, 0), w - 1), sy = min(max(y0 - R + iy, 0), h - 1);
        in[i] = __ldg(src + (size_t)sy * w + sx);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < IH * ST_TW; i += ST_THREADS) {
        const int row = i / ST_TW, col = i - row * ST_TW;
        const float4* p = in + row * IW + col;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int t = 0; t <= 2 * R; ++t) acc = fma4s(K.g[t], p[t], acc);
        mid[i] = acc;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < ST_PER_THREAD; ++k) {
        const int i = threadIdx.x + k * ST_THREADS;
        const int row = i / ST_TW, col = i - row * ST_TW;
        const float4* p = mid + row * ST_TW + col;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int t = 0; t <= 2 * R; ++t) acc = fma4s(K.g[t], p[t * ST_TW], acc);
        out[k] = make_float4(acc.x * K.scale, acc.y * K.scale, acc.z * K.scale, acc.w * K.scale);
    }
}

__device__ __forceinline__ void tile_load(const float4* __restrict__ src, int w, int h, int x0, int y0, float4 out[ST_PER_THREAD])
{
#pragma unroll
    for (int k = 0; k < ST_PER_THREAD; ++k) {
        const int i = threadIdx.x + k * ST_THREADS;
        const int row = i / ST_TW, col = i - row * ST_TW;
        const int x = min(x0 + col, w - 1), y = min(y0 + row, h - 1);
        out[k] = __ldg(src + (size_t)y * w + x);
    }
}

__global__ void __launch_bounds__(ST_THREADS) k_stack_f32(float4* __restrict__ dst, const __grid_constant__ StParams P)
{
    extern __shared__ float4 smem[];
    float4* in = smem;
    float4* mid = smem + (ST_TH + 2 * ST_MAX_R) * (ST_TW + 2 * ST_MAX_R);
    const int x0 = blockIdx.x * ST_TW, y0 = blockIdx.y * ST_TH;

    float4 acc[ST_PER_THREAD];
#pragma unroll
    for (int k = 0; k < ST_PER_THREAD; ++k) acc[k] = P.background;

    for (int l = 0; l < P.n_layers; ++l) {
        const StLayer& L = P.layer[l];
        float4 a[ST_PER_THREAD];
        if (L.ka >= 0) tile_blur(L.a, P.w, P.h, x0, y0, P.k[L.ka], in, mid, a);
        else tile_load(L.a, P.w, P.h, x0, y0, a);
        if (L.b) {
            float4 b[ST_PER_THREAD];
            if (L.kb >= 0) tile_blur(L.b, P.w, P.h, x0, y0, P.k[L.kb], in, mid, b);
            else tile_load(L.b, P.w, P.h, x0, y0, b);
            // Dissolve: add(mul(from, iv), mul(to, v))
#pragma unroll
            for (int k = 0; k < ST_PER_THREAD; ++k) {
                a[k].x = __fadd_rn(__fmul_rn(a[k].x, L.iv), __fmul_rn(b[k].x, L.v));
                a[k].y = __fadd_rn(__fmul_rn(a[k].y, L.iv), __fmul_rn(b[k].y, L.v));
                a[k].z = __fadd_rn(__fmul_rn(a[k].z, L.iv), __fmul_rn(b[k].z, L.v));
                a[k].w = __fadd_rn(__fmul_rn(a[k].w, L.iv), __fmul_rn(b[k].w, L.v));
            }
        }
        // CompNode: premult(fg) then over(fg, acc)
#pragma unroll
        for (int k = 0; k < ST_PER_THREAD; ++k) {
            float4 f = a[k];
            if (f.w != 1.0f) {
                f.x = __fmul_rn(f.x, f.w);
                f.y = __fmul_rn(f.y, f.w);
                f.z = __fmul_rn(f.z, f.w);
            }
            const float oma = __fsub_rn(1.0f, fminf(fmaxf(f.w, 0.0f), 1.0f));
            acc[k].x = __fadd_rn(f.x, __fmul_rn(oma, acc[k].x));
            acc[k].y = __fadd_rn(f.y, __fmul_rn(oma, acc[k].y));
            acc[k].z = __fadd_rn(f.z, __fmul_rn(oma, acc[k].z));
            acc[k].w = __fadd_rn(f.w, __fmul_rn(oma, acc[k].w));
        }
    }
#pragma unroll
    for (int k = 0; k < ST_PER_THREAD; ++k) {
        const int i = threadIdx.x + k * ST_THREADS;
        const int row = i / ST_TW, col = i - row * ST_TW;
        const int x = x0 + col, y = y0 + row;
        if (x < P.w && y < P.h) dst[(size_t)y * P.w + x] = acc[k];
    }
}

// same construction as tc_conv.cu build_params (OIIO make_kernel + convolve normalisation)
static float st_fast_exp(float xin)
{
    float x = xin * (float)(1.0 / M_LN2);
    x = x < -126.0f ? -126.0f : (x > 126.0f ? 126.0f : x);
    int m = (int)x;
    x -= (float)m;
    x = 1.0f - (1.0f - x);
    volatile float r = 1.33336498402e-3f;
    const float c[5] = { 9.810352697968e-3f, 5.551834031939e-2f, 0.2401793301105f, 0.693144857883f, 1.0f };
    for (int i = 0; i < 5; ++i) {
        r = x * r;
        r = r + c[i];
    }
    float rr = r;
    uint32_t bits;
    memcpy(&bits, &rr, 4);
    bits += ((uint32_t)m << 23);
    memcpy(&rr, &bits, 4);
    return rr;
}
static int st_build_kernel(float width, StKernel* K)
{
    int n = (int)ceilf(width);
    if (n < 1) n = 1;
    if (!(n & 1)) ++n;
    const int r = n / 2;
    std::vector<float> g(n);
    const float radinv = 2.0f / width;
    for (int i = 0; i < n; ++i) {
        const float u = fabsf((float)(i - r) * radinv);
        g[i] = u < 1.0f ? st_fast_exp(-2.0f * (u * u)) : 0.0f;
    }
    volatile float sum = 0.0f;
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) sum = sum + g[i] * g[j];
    volatile float ksum = 0.0f;
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) ksum = ksum + (g[i] * g[j]) / sum;
    int live = r;
    while (live > 0 && g[r - live] == 0.0f && g[r + live] == 0.0f) --live;
    if (live > ST_MAX_R) return fail(TC_ERR_UNSUPPORTED, "tc_stack_f32: blur radius %g needs %d taps per side (max %d)", width, live, ST_MAX_R);
    K->R = live;
    K->scale = (1.0f / ksum) / sum;
    memset(K->g, 0, sizeof(K->g));
    for (int i = -live; i <= live; ++i) K->g[i + live] = g[r + i];
    return TC_OK;
}

} // namespace tc

using namespace tc;

extern "C" int tc_stack_f32(const tc_image* dst, const float background[4], const tc_stack_layer* layers, int n_layers, tc_stream s)
{
    int v;
    if ((v = validate(dst, "tc_stack_f32 dst")) != TC_OK) return v;
    if (dst->type != TC_F32) return fail(TC_ERR_UNSUPPORTED, "tc_stack_f32: destination must be float32");
    if (!background || (n_layers > 0 && !layers)) return fail(TC_ERR_INVALID, "tc_stack_f32: null argument");
    if (n_layers < 0 || n_layers > ST_MAX_LAYERS) return fail(TC_ERR_UNSUPPORTED, "tc_stack_f32: %d layers (max %d)", n_layers, ST_MAX_LAYERS);
    StParams P;
    memset(&P, 0, sizeof(P));
    P.w = dst->w;
    P.h = dst->h;
    P.n_layers = n_layers;
    P.background = make_float4(background[0], background[1], background[2], background[3]);
    float radii[ST_MAX_KERNELS];
    int nk = 0;
    auto kernel_of = [&](float radius, int* idx) -> int {
        *idx = -1;
        if (!(radius > 0.0f)) return TC_OK;
        for (int i = 0; i < nk; ++i)
            if (radii[i] == radius) {
                *idx = i;
                return TC_OK;
            }
        if (nk == ST_MAX_KERNELS) return fail(TC_ERR_UNSUPPORTED, "tc_stack_f32: more than %d distinct blur radii", ST_MAX_KERNELS);
        int st = st_build_kernel(radius, &P.k[nk]);
        if (st != TC_OK) return st;
        radii[nk] = radius;
        *idx = nk++;
        return TC_OK;
    };
    for (int l = 0; l < n_layers; ++l) {
        if (!layers[l].src_a) return fail(TC_ERR_INVALID, "tc_stack_f32: layer %d has no source", l);
        P.layer[l].a = (const float4*)layers[l].src_a;
        P.layer[l].b = (const float4*)layers[l].src_b;
        if ((v = kernel_of(layers[l].radius_a, &P.layer[l].ka)) != TC_OK) return v;
        P.layer[l].kb = -1;
        if (layers[l].src_b && (v = kernel_of(layers[l].radius_b, &P.layer[l].kb)) != TC_OK) return v;
        P.layer[l].v = (float)layers[l].value;
        P.layer[l].iv = (float)(1.0 - layers[l].value);
    }
    const size_t smem = (size_t)((ST_TH + 2 * ST_MAX_R) * (ST_TW + 2 * ST_MAX_R) + (ST_TH + 2 * ST_MAX_R) * ST_TW) * sizeof(float4);
    static bool configured = false;
    if (!configured) {
        TC_CUDA(cudaFuncSetAttribute(k_stack_f32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    dim3 grid((dst->w + ST_TW - 1) / ST_TW, (dst->h + ST_TH - 1) / ST_TH);
    k_stack_f32<<<grid, ST_THREADS, smem, (cudaStream_t)s>>>((float4*)dst->data, P);
    count_launch();
    return check_cuda(cudaGetLastError(), "k_stack_f32 launch");
}
