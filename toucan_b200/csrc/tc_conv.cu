// tc_conv.cu -- toucan:Blur and toucan:UnsharpMask: the reference's 2-D Gaussian
// convolve (O(k^2) taps per pixel, plugins/FilterPlugin.cpp:188-202 / :607-618) as a
// separable shared-memory tiled stencil, one read of the source and one write of the
// result per pixel.  Edge handling = clamp (OIIO convolve uses WrapClamp), so border
// tiles are staged with clamped coordinates rather than zero-filling bulk copies.
//
// Tolerance op (<= 1e-5 abs / +-1 LSB), so FMA contraction is allowed in this file.
#include <cmath>
#include <cstring>
#include <vector>

#include "tc_tile.cuh"

namespace tc {

constexpr int CONV_MAX_R = 24;   // live taps up to 49 (radius parameter up to ~50)
constexpr int CONV_TW = 32;
constexpr int CONV_TH = 32;
constexpr int CONV_THREADS = 256;

struct ConvParams {
    int R;           // live half-width: taps -R..R
    float scale;     // (1/sum K2D) * (1/sum g g): final normalisation
    float contrast;  // unsharp only
    float threshold; // unsharp only
    float g[2 * CONV_MAX_R + 1];
};

__device__ __forceinline__ float4 fma4(float w, float4 p, float4 a)
{
    return make_float4(fmaf(w, p.x, a.x), fmaf(w, p.y, a.y), fmaf(w, p.z, a.z), fmaf(w, p.w, a.w));
}

// Diff = src - Blurry; threshold; *contrast; dst = src + Diff  (OIIO unsharp_mask)
__device__ __forceinline__ float4 unsharp_px(float4 sp, float4 o, float contrast, float threshold)
{
    float d[4] = { sp.x - o.x, sp.y - o.y, sp.z - o.z, sp.w - o.w };
    const float s4[4] = { sp.x, sp.y, sp.z, sp.w };
    float r[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        float dd = d[c];
        if (threshold > 0.0f && fabsf(dd) < threshold) dd = 0.0f;
        dd = __fmul_rn(dd, contrast);
        r[c] = __fadd_rn(s4[c], dd);
    }
    return make_float4(r[0], r[1], r[2], r[3]);
}

// Stage the source tile (+R halo, coordinates clamped: OIIO convolve uses WrapClamp) as float: thread = one staged
// column, STAGE_ROWS rows per pass, six loads in flight per thread before the first conversion.
template <int R, int TH, int TYPE>
__device__ __forceinline__ void gauss_stage(const Img& src, int x0, int y0, float4* in)
{
    using T = StTile<R, TH, 1>;
    const int tid = threadIdx.x;
    if (tid < T::STAGE_ROWS * T::IW) {
        const int ry = tid / T::IW, ix = tid - ry * T::IW;
        const int sx = min(max(x0 - R + ix, 0), src.w - 1);
        constexpr int BATCH = 6;
#pragma unroll
        for (int u0 = 0; u0 < T::STAGE_N; u0 += BATCH) {
            float4 v[BATCH];
#pragma unroll
            for (int u = 0; u < BATCH; ++u) {
                const int iy = ry + (u0 + u) * T::STAGE_ROWS;
                if (u0 + u < T::STAGE_N && iy < T::IH)
                    v[u] = load_px(src.p, TYPE, (size_t)min(max(y0 - R + iy, 0), src.h - 1) * src.w + sx);
            }
#pragma unroll
            for (int u = 0; u < BATCH; ++u) {
                const int iy = ry + (u0 + u) * T::STAGE_ROWS;
                if (u0 + u < T::STAGE_N && iy < T::IH) in[iy * T::IWP + ix] = v[u];
            }
        }
    }
}

// Register-blocked tile kernel (tc_tile.cuh), any storage type in/out, R <= 12.
// One CTA per 64 x 24 output tile: stage (typed load -> float, clamped) -> rows -> columns ->
// [unsharp epilogue] -> typed store.  P.g carries sqrt(scale).
template <int R, bool UNSHARP, int OCC>
__global__ void __launch_bounds__(ST_THREADS, OCC) k_gauss_tile(const Img src, const Img dst, const __grid_constant__ ConvParams P)
{
    constexpr int TH = 24;
    using T = StTile<R, TH, 1>;
    extern __shared__ float4 smem[];
    float4* in = smem;
    float4* mid = smem + T::IN_ELEMS;
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * ST_TW, y0 = blockIdx.y * TH;

    // one load path per storage type, chosen once: with the type switch inside the batch every typed load was
    // followed by its conversion, which waits for it -- six exposed load latencies per batch instead of one
    // (half / 8-bit sources ran 40 % slower than float ones)
    switch (src.type) {
    case TC_U8: gauss_stage<R, TH, TC_U8>(src, x0, y0, in); break;
    case TC_U16: gauss_stage<R, TH, TC_U16>(src, x0, y0, in); break;
    case TC_F16: gauss_stage<R, TH, TC_F16>(src, x0, y0, in); break;
    default: gauss_stage<R, TH, TC_F32>(src, x0, y0, in); break;
    }
    float g[2 * R + 1];
#pragma unroll
    for (int t = 0; t <= 2 * R; ++t) {
        const int s = t - R + P.R;
        g[t] = (s >= 0 && s <= 2 * P.R) ? P.g[s] : 0.0f;
    }
    __syncthreads();
    tile_rows<R, TH, true>(in, mid, g);
    __syncthreads();
    float4 f[T::PX];
    tile_cols<R, TH>(mid, g, f);

    const int col = tid % ST_TW, seg = tid / ST_TW;
    const int x = x0 + col, y = y0 + seg * T::PX;
    const bool same = src.w == dst.w && src.h == dst.h;
    if (x < dst.w) {
#pragma unroll
        for (int k = 0; k < T::PX; ++k) {
            if (y + k < dst.h) {
                float4 o = f[k];
                // the unfiltered pixel is still in the staged tile (same-sized source: the clamped tile equals the source there)
                if (UNSHARP)
                    o = unsharp_px(same ? in[(seg * T::PX + k + R) * T::IWP + col + R] : load_black(src, x, y + k), o, P.contrast, P.threshold);
                store_px(dst.p, dst.type, (size_t)(y + k) * dst.w + x, o);
            }
        }
    }
}

// Generic fallback for very wide kernels (R > 12): 32 x 32 tiles, one tap per shared-memory read.
template <bool UNSHARP>
__global__ void __launch_bounds__(CONV_THREADS) k_gauss_wide(const Img src, const Img dst, const __grid_constant__ ConvParams P)
{
    extern __shared__ float4 smem[];
    const int R = P.R;
    const int IW = CONV_TW + 2 * R, IH = CONV_TH + 2 * R;
    float4* in = smem;            // [IH][IW]
    float4* mid = smem + IH * IW; // [IH][CONV_TW]
    const int x0 = blockIdx.x * CONV_TW, y0 = blockIdx.y * CONV_TH;

    for (int i = threadIdx.x; i < IH * IW; i += CONV_THREADS) {
        const int iy = i / IW, ix = i - iy * IW;
        in[i] = load_clamp(src, x0 - R + ix, y0 - R + iy);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < IH * CONV_TW; i += CONV_THREADS) {
        const int row = i / CONV_TW, col = i - row * CONV_TW;
        const float4* p = in + row * IW + col;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int t = 0; t <= 2 * R; ++t) acc = fma4(P.g[t], p[t], acc);
        mid[i] = acc;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < CONV_TH * CONV_TW; i += CONV_THREADS) {
        const int row = i / CONV_TW, col = i - row * CONV_TW;
        const int x = x0 + col, y = y0 + row;
        if (x >= dst.w || y >= dst.h) continue;
        const float4* p = mid + row * CONV_TW + col;
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int t = 0; t <= 2 * R; ++t) o = fma4(P.g[t], p[t * CONV_TW], o);
        if (UNSHARP) o = unsharp_px(load_black(src, x, y), o, P.contrast, P.threshold);
        store_px(dst.p, dst.type, (size_t)y * dst.w + x, o);
    }
}

// ---- host: kernel weights exactly as OIIO builds them ---------------------------------
// OIIO fmath.h fast_exp2 / fast_exp (no FMA on the reference build).
static float fast_exp2_host(float xval)
{
    float x = xval < -126.0f ? -126.0f : (xval > 126.0f ? 126.0f : xval);
    int m = (int)x;
    x -= (float)m;
    x = 1.0f - (1.0f - x);
    volatile float r = 1.33336498402e-3f; // volatile: keep every step rounded to float, unfused
    r = x * r;
    r = r + 9.810352697968e-3f;
    r = x * r;
    r = r + 5.551834031939e-2f;
    r = x * r;
    r = r + 0.2401793301105f;
    r = x * r;
    r = r + 0.693144857883f;
    r = x * r;
    r = r + 1.0f;
    float rr = r;
    uint32_t bits;
    memcpy(&bits, &rr, 4);
    bits += ((uint32_t)m << 23);
    memcpy(&rr, &bits, 4);
    return rr;
}
static float fast_exp_host(float x) { return fast_exp2_host(x * (float)(1.0 / M_LN2)); }

// make_kernel("gaussian", width, width) + convolve(normalize=true): separable factors.
static int build_params(float width, ConvParams* P)
{
    if (!(width > 0.0f)) return fail(TC_ERR_INVALID, "gaussian kernel width must be > 0 (got %g)", width);
    int n = (int)ceilf(width);
    if (n < 1) n = 1;
    if (!(n & 1)) ++n;
    const int r = n / 2;
    std::vector<float> g(n);
    const float radinv = 2.0f / width;
    for (int i = 0; i < n; ++i) {
        const float u = fabsf((float)(i - r) * radinv);
        g[i] = u < 1.0f ? fast_exp_host(-2.0f * (u * u)) : 0.0f;
    }
    // 2-D kernel sum in the reference's order (make_kernel normalisation), then the
    // convolve() normalisation over the normalised kernel.
    volatile float sum = 0.0f;
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) sum = sum + g[i] * g[j];
    volatile float ksum = 0.0f;
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) ksum = ksum + (g[i] * g[j]) / sum;
    // trim zero-weight edge taps
    int live = r;
    while (live > 0 && g[r - live] == 0.0f && g[r + live] == 0.0f) --live;
    if (live > CONV_MAX_R)
        return fail(TC_ERR_UNSUPPORTED, "gaussian kernel of width %g needs %d taps per side (max %d)", width, live, CONV_MAX_R);
    P->R = live;
    P->scale = (1.0f / ksum) / sum;
    // each 1-D pass carries sqrt(scale): the two passes multiply out to the normalised 2-D kernel
    const float root = sqrtf(P->scale);
    memset(P->g, 0, sizeof(P->g));
    for (int i = -live; i <= live; ++i) P->g[i + live] = g[r + i] * root;
    return TC_OK;
}

template <int R, bool UNSHARP>
static int launch_tile(const tc_image* dst, const tc_image* src, const ConvParams& P, cudaStream_t st)
{
    using T = StTile<R, 24, 1>;
    constexpr int OCC = T::SMEM <= 74 * 1024 ? 3 : (T::SMEM <= 112 * 1024 ? 2 : 1);
    int st_attr = ensure_dynamic_smem((const void*)k_gauss_tile<R, UNSHARP, OCC>, T::SMEM);
    if (st_attr != TC_OK) return st_attr;
    dim3 grid((dst->w + ST_TW - 1) / ST_TW, (dst->h + 23) / 24);
    k_gauss_tile<R, UNSHARP, OCC><<<grid, ST_THREADS, T::SMEM, st>>>(view(src), view(dst), P);
    count_launch();
    return check_cuda(cudaGetLastError(), "k_gauss_tile launch");
}

template <bool UNSHARP>
static int launch_conv_t(const tc_image* dst, const tc_image* src, const ConvParams& P, cudaStream_t st)
{
    if (P.R <= 2) return launch_tile<2, UNSHARP>(dst, src, P, st);
    if (P.R <= 4) return launch_tile<4, UNSHARP>(dst, src, P, st);
    if (P.R <= 6) return launch_tile<6, UNSHARP>(dst, src, P, st);
    if (P.R <= 9) return launch_tile<9, UNSHARP>(dst, src, P, st);
    if (P.R <= 12) return launch_tile<12, UNSHARP>(dst, src, P, st);
    const int IW = CONV_TW + 2 * P.R, IH = CONV_TH + 2 * P.R;
    const size_t smem = (size_t)(IH * IW + IH * CONV_TW) * sizeof(float4);
    int st_attr = ensure_dynamic_smem((const void*)k_gauss_wide<UNSHARP>, smem);
    if (st_attr != TC_OK) return st_attr;
    dim3 grid((dst->w + CONV_TW - 1) / CONV_TW, (dst->h + CONV_TH - 1) / CONV_TH);
    k_gauss_wide<UNSHARP><<<grid, CONV_THREADS, smem, st>>>(view(src), view(dst), P);
    count_launch();
    return check_cuda(cudaGetLastError(), "k_gauss_wide launch");
}

static int launch_conv(const tc_image* dst, const tc_image* src, const ConvParams& P, bool unsharp, tc_stream s)
{
    return unsharp ? launch_conv_t<true>(dst, src, P, (cudaStream_t)s) : launch_conv_t<false>(dst, src, P, (cudaStream_t)s);
}

} // namespace tc

using namespace tc;

extern "C" {

int tc_blur(const tc_image* dst, const tc_image* src, float radius, tc_stream s)
{
    int v;
    if ((v = validate(dst, "tc_blur dst")) != TC_OK) return v;
    if ((v = validate(src, "tc_blur src")) != TC_OK) return v;
    ConvParams P;
    if ((v = build_params(radius, &P)) != TC_OK) return v;
    P.contrast = 0.f;
    P.threshold = 0.f;
    return launch_conv(dst, src, P, false, s);
}

int tc_unsharp_mask(const tc_image* dst, const tc_image* src, const char* kernel, float width, float contrast,
                    float threshold, tc_stream s)
{
    int v;
    if ((v = validate(dst, "tc_unsharp_mask dst")) != TC_OK) return v;
    if ((v = validate(src, "tc_unsharp_mask src")) != TC_OK) return v;
    if (!kernel || strcmp(kernel, "gaussian") != 0)
        return fail(TC_ERR_UNSUPPORTED, "tc_unsharp_mask: kernel '%s' (only 'gaussian' is in scope)", kernel ? kernel : "(null)");
    ConvParams P;
    if ((v = build_params(width, &P)) != TC_OK) return v;
    P.contrast = contrast;
    P.threshold = threshold;
    return launch_conv(dst, src, P, true, s);
}

} // extern "C"
