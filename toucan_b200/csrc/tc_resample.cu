// tc_resample.cu -- toucan:Resize and toucan:Rotate (and the fit-resize used by
// CompNode / Dissolve): OIIO's filtered resample restated as gathers.
//
//   resize: separable weights (blackman-harris when enlarging, lanczos3 otherwise) are
//           built once per launch by a tiny per-axis kernel into tap tables, then a row
//           pass and a column pass evaluate  sum_j wy_j sum_i wx_i src(clamp)  with
//           2(2r+1) taps per pixel instead of (2r+1)^2
//           (plugins/TransformPlugin.cpp:374-379 -> OIIO resize_, separable branch).
//   rotate: inverse-maps each pixel centre and filters a lanczos3 footprint with black
//           outside the source (TransformPlugin.cpp:462-467 -> OIIO rotate/warp_/filtered_sample).
//
// Compiled with -fmad=false so the per-tap arithmetic follows the reference's unfused
// float sequence; lanczos3 uses OIIO's fast_sinpi polynomial (max abs error 9.2e-4), NOT
// sinpif -- the difference is far above the 1e-5 tolerance.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <tuple>
#include <type_traits>

#include "tc_common.cuh"
#include "tc_tile.cuh"
#include "tc_tma.cuh"

namespace tc {

__device__ __forceinline__ float fast_sinpi(float x)
{
    const float z = __fsub_rn(x, __fsub_rn(__fadd_rn(x, 25165824.0f), 25165824.0f));
    const float y = __fsub_rn(z, __fmul_rn(z, fabsf(z)));
    const float Q = 3.10396624f;
    const float P = 3.584135056f;
    return __fmul_rn(y, __fadd_rn(Q, __fmul_rn(P, fabsf(y))));
}
__device__ __forceinline__ float lanczos3(float x)
{
    const float a = 3.0f;
    const float ainv = 1.0f / 3.0f;
    const float m_pi = 3.14159265358979323846f;
    x = fabsf(x);
    if (x > a) return 0.0f;
    if (x < 0.0001f) return 1.0f;
    const float d = __fmul_rn(__fmul_rn(x, x), __fmul_rn(m_pi, m_pi));
    return __fmul_rn(__fmul_rn(__fdiv_rn(a, d), fast_sinpi(x)), fast_sinpi(__fmul_rn(x, ainv)));
}
__device__ __forceinline__ float bh1d(float x)
{
    if (x < -1.0f || x > 1.0f) return 0.0f;
    x = __fmul_rn(__fadd_rn(x, 1.0f), 0.5f);
    const float A0 = 0.35875f, A1 = -0.48829f, A2 = 0.14128f, A3 = -0.01168f;
    const float m_pi = 3.14159265358979323846f;
    float r = __fadd_rn(A0, __fmul_rn(A1, cosf(__fmul_rn(__fmul_rn(2.f, m_pi), x))));
    r = __fadd_rn(r, __fmul_rn(A2, cosf(__fmul_rn(__fmul_rn(4.f, m_pi), x))));
    r = __fadd_rn(r, __fmul_rn(A3, cosf(__fmul_rn(__fmul_rn(6.f, m_pi), x))));
    return r;
}
// id 0 = lanczos3 (argument scale 6/width), 1 = blackman-harris (2/width)
__device__ __forceinline__ float filt1d(int id, float scale, float x)
{
    return id == 0 ? lanczos3(__fmul_rn(x, scale)) : bh1d(__fmul_rn(x, scale));
}

struct AxisParams {
    int dn, sn;    // destination / source extent
    int id;        // filter id
    float scale;   // filter argument scale
    float ratio;   // dn / sn
    int radi, taps;
};

// One thread per destination index: first source tap + `taps` normalised weights.
__global__ void k_resize_axis(const AxisParams A, int* __restrict__ first, float* __restrict__ wts, float* __restrict__ total)
{
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= A.dn) return;
    const float dpix = __fdiv_rn(1.0f, (float)A.dn);
    const float s = __fmul_rn(__fadd_rn((float)d, 0.5f), dpix);
    const float sf = __fadd_rn(0.0f, __fmul_rn(s, (float)A.sn));
    const float fl = floorf(sf);
    const int si = (int)fl;
    const float frac = __fsub_rn(sf, fl);
    float tot = 0.0f;
    float* w = wts + (size_t)d * A.taps;
    for (int i = 0; i < A.taps; ++i) {
        const float v = filt1d(A.id, A.scale, __fmul_rn(A.ratio, __fsub_rn((float)(i - A.radi), __fsub_rn(frac, 0.5f))));
        w[i] = v;
        tot = __fadd_rn(tot, v);
    }
    float resum = 0.0f;
    if (tot != 0.0f)
        for (int i = 0; i < A.taps; ++i) {
            w[i] = __fdiv_rn(w[i], tot);
            resum = __fadd_rn(resum, w[i]);
        }
    else
        for (int i = 0; i < A.taps; ++i) resum = __fadd_rn(resum, w[i]);
    first[d] = si - A.radi;
    // x axis consumers test the re-summed normalised weights, y axis the raw total
    total[d * 2 + 0] = tot;
    total[d * 2 + 1] = resum;
}

// Separable evaluation of  sum_j wy_j sum_i wx_i src(clamp)  in two passes through a float32
// scratch image (the reference loops all (2r+1)^2 taps per pixel with the same separable weights;
// the two-pass order differs from it by float rounding only, far inside the 1e-5 tolerance):
//   k_resize_rows:  tmp(x, sy) = sum_i wx_i(x) * src(clamp(fx(x) + i), sy)      dw x sh
//   k_resize_cols:  dst(x, y)  = sum_j wy_j(y) * tmp(x, clamp(fy(y) + j))       dw x dh
__global__ void __launch_bounds__(256) k_resize_rows(const Img src, float4* __restrict__ tmp, int dw, const int* __restrict__ xf,
                                                      const float* __restrict__ xw, const float* __restrict__ xt, int xtaps)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int sy = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dw || sy >= src.h) return;
    float4 pel = make_float4(0.f, 0.f, 0.f, 0.f);
    if (xt[x * 2 + 1] != 0.0f) {
        const float* wxp = xw + (size_t)x * xtaps;
        const int fx = xf[x];
        const size_t row = (size_t)sy * src.w;
        for (int i = 0; i < xtaps; ++i) {
            const float w = wxp[i];
            if (w != 0.0f) {
                const float4 p = load_px(src.p, src.type, row + min(max(fx + i, 0), src.w - 1));
                pel.x = fmaf(w, p.x, pel.x);
                pel.y = fmaf(w, p.y, pel.y);
                pel.z = fmaf(w, p.z, pel.z);
                pel.w = fmaf(w, p.w, pel.w);
            }
        }
    }
    tmp[(size_t)sy * dw + x] = pel;
}

__global__ void __launch_bounds__(256) k_resize_cols(const float4* __restrict__ tmp, int sh, const Img dst, const int* __restrict__ yf,
                                                      const float* __restrict__ yw, const float* __restrict__ yt, int ytaps)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dst.w || y >= dst.h) return;
    float4 pel = make_float4(0.f, 0.f, 0.f, 0.f);
    if (yt[y * 2 + 0] != 0.0f) {
        const float* wyp = yw + (size_t)y * ytaps;
        const int fy = yf[y];
        for (int j = 0; j < ytaps; ++j) {
            const float w = wyp[j];
            if (w != 0.0f) {
                const float4 p = __ldg(tmp + (size_t)min(max(fy + j, 0), sh - 1) * dst.w + x);
                pel.x = fmaf(w, p.x, pel.x);
                pel.y = fmaf(w, p.y, pel.y);
                pel.z = fmaf(w, p.z, pel.z);
                pel.w = fmaf(w, p.w, pel.w);
            }
        }
    }
    store_px(dst.p, dst.type, (size_t)y * dst.w + x, pel);
}

// Very large reductions (thousands of taps per pixel: thumbnails of big frames): the sum in the REFERENCE's order -- rows outer,
// columns inner, the weight product rounded, multiply and add unfused (OIIO resize_, separable branch) -- because at 20 000+
// taps the rounding of any other order (1.5e-5 measured at 402 -> 1 pixels) exceeds the 1e-5 bar.  One thread per destination
// pixel; only dispatched while destination pixels x taps stays small.
template <int TYPE>
__global__ void __launch_bounds__(128) k_resize_direct(const Img src_, const Img dst, const int* __restrict__ xf, const float* __restrict__ xw,
                                                        const float* __restrict__ xt, const int* __restrict__ yf, const float* __restrict__ yw,
                                                        const float* __restrict__ yt, int xtaps, int ytaps)
{
    const Img src { src_.p, src_.w, src_.h, TYPE };
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= dst.w * dst.h) return;
    const int y = i / dst.w, x = i - y * dst.w;
    float4 pel = make_float4(0.f, 0.f, 0.f, 0.f);
    if (xt[x * 2 + 1] != 0.0f) {
        const float* wxp = xw + (size_t)x * xtaps;
        const float* wyp = yw + (size_t)y * ytaps;
        const int fx = xf[x], fy = yf[y];
        for (int j = 0; j < ytaps; ++j) {
            const float wy = wyp[j];
            if (wy == 0.0f) continue;
            const size_t row = (size_t)min(max(fy + j, 0), src.h - 1) * src.w;
            for (int k = 0; k < xtaps; ++k) {
                const float w = __fmul_rn(wy, wxp[k]);
                if (w != 0.0f) {
                    const float4 p = load_px(src.p, TYPE, row + min(max(fx + k, 0), src.w - 1));
                    pel.x = __fadd_rn(pel.x, __fmul_rn(w, p.x));
                    pel.y = __fadd_rn(pel.y, __fmul_rn(w, p.y));
                    pel.z = __fadd_rn(pel.z, __fmul_rn(w, p.z));
                    pel.w = __fadd_rn(pel.w, __fmul_rn(w, p.w));
                }
            }
        }
    }
    if (yt[y * 2 + 0] == 0.0f) pel = make_float4(0.f, 0.f, 0.f, 0.f);
    store_px(dst.p, dst.type, (size_t)i, pel);
}

// Row pass of the two-pass path through the same kind of shared tile, for footprints the march cannot hold (more
// than 32 taps per axis or 128 staged columns: reductions below ~0.3).  A CTA owns 32 source rows x `xw` destination columns; the rows' common
// source span is staged once (coalesced), then lane = row, warp = destination column, weights warp-uniform -- instead of
// every thread gathering its `taps` source pixels with a stride of 1/ratio pixels through L1.
constexpr int RR_ROWS = 32, RR_THREADS = 256;

struct RowsTile { int cols; int xw; int xtaps; }; // staged columns (allocation), destination columns per CTA

template <int TYPE>
__global__ void __launch_bounds__(RR_THREADS) k_resize_rows_tile(const Img src_, float4* __restrict__ tmp, int dw, const int* __restrict__ xf,
                                                                  const float* __restrict__ xw, const float* __restrict__ xt, const RowsTile T)
{
    extern __shared__ float4 rr_smem[];
    const Img src { src_.p, src_.w, src_.h, TYPE };
    const int pitch = T.cols | 1;
    float4* S = rr_smem;                                              // [RR_ROWS][pitch]
    float* ws = reinterpret_cast<float*>(S + (size_t)RR_ROWS * pitch); // [xw][xtaps]
    int* fs = reinterpret_cast<int*>(ws + T.xw * T.xtaps);            // [xw]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int WARPS = RR_THREADS / 32;
    const int x0 = blockIdx.x * T.xw, y0 = blockIdx.y * RR_ROWS;
    const int nx = min(T.xw, dw - x0), ny = min(RR_ROWS, src.h - y0);
    const int fx0 = xf[x0];
    const int cols = min(xf[x0 + nx - 1] + T.xtaps - fx0, T.cols);
    for (int i = tid; i < nx * T.xtaps; i += RR_THREADS) {
        const int x = i / T.xtaps, t = i - x * T.xtaps;
        ws[i] = xt[(x0 + x) * 2 + 1] != 0.0f ? xw[(size_t)(x0 + x) * T.xtaps + t] : 0.0f;
    }
    if (tid < nx) fs[tid] = min(xf[x0 + tid] - fx0, T.cols - T.xtaps);
    // stage: a warp per source row, eight loads in flight per thread
    for (int r = warp; r < ny; r += 2 * WARPS) {
        const int rb = min(r + WARPS, ny - 1);
        const size_t ra = (size_t)(y0 + r) * src.w, rbase = (size_t)(y0 + rb) * src.w;
        for (int c0 = 0; c0 < cols; c0 += 128) {
            float4 va[4], vb[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int sx = min(max(fx0 + min(c0 + k * 32 + lane, cols - 1), 0), src.w - 1);
                va[k] = load_px(src.p, TYPE, ra + sx);
                vb[k] = load_px(src.p, TYPE, rbase + sx);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int c = c0 + k * 32 + lane;
                if (c < cols) {
                    S[(size_t)r * pitch + c] = va[k];
                    S[(size_t)rb * pitch + c] = vb[k];
                }
            }
        }
    }
    __syncthreads();
    const float4* row = S + (size_t)min(lane, ny - 1) * pitch;
    for (int x = warp; x < nx; x += WARPS) {
        const float4* p = row + fs[x];
        const float* w = ws + x * T.xtaps;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
        for (int t = 0; t < T.xtaps; ++t) {
            const float wt = w[t];
            if (wt != 0.0f) acc = fma4p(wt, p[t], acc);
        }
        if (lane < ny) tmp[(size_t)(y0 + lane) * dw + x0 + x] = acc;
    }
}

// The two passes fused as a MARCH down the frame (the north-star layout for Resize: a filtered-resample gather through
// shared-memory tiles; round 1's 32 x 10 output tile re-staged every source row ~3 times at 1/2).  A CTA owns a strip of 32 destination columns and a chunk of
// destination rows and walks its source rows in bands of 32:
//   stage    the band's source span (32 rows x the strip's footprint columns, edge-clamped, float) -- each source row
//            is staged ONCE per strip (plus ytaps warm-up rows per chunk);
//   row pass lane = source row (odd 16-byte row pitch: conflict-free), warp = a block of 4 destination columns whose
//            common source span is read once: 4 outputs from span = 3 steps + taps reads (19-20 at 1/2 instead of 52).
//            The per-block weights are expanded to a span x 4 matrix in shared memory once per CTA (zero where a source
//            pixel is outside an output's window), so the loop has no per-tap index arithmetic: one 128-bit pixel read,
//            one broadcast 128-bit weight read, 8 FFMA2.  Results go into a ring of 64 row slots, Mt[column][slot];
//   column pass lane = destination column (pitch 65: conflict-free), warp = a block of 4 destination rows, same scheme
//            with the weights expanded per block by the warp; runs for every destination row whose window is complete.
// Same weights and the same ascending-tap fused accumulation as k_resize_rows / k_resize_cols, so frames are identical
// to theirs for finite pixels (the expanded matrix adds exact zeros).
constexpr int RM_W = 32, RM_BAND = 32, RM_THREADS = 256, RM_B = 4, RM_SPAN = 48;
constexpr int RM_WARPS = RM_THREADS / 32, RM_XBLK = RM_W / RM_B, RM_MAXBANDS = RM_THREADS;

struct ResizeMarch {
    int sw;       // staged columns (allocation)
    int ch;       // destination rows per chunk
    int xtaps, ytaps;
    int ring;     // row slots of the ring: 48 when ytaps <= 16 (a band of 32 new rows + the rows still needed), else 64
    int wspan;    // row pitch of the expanded weight tables: the longest block span, <= RM_SPAN
};

static size_t resize_march_smem(int sw, int ring, int wspan)
{
    return ((size_t)RM_BAND * (sw | 1) + (size_t)RM_W * (ring + 1) + (size_t)(RM_XBLK + RM_WARPS) * wspan) * sizeof(float4);
}

// a pixel as stored (1 / 2 / 2 / 4 registers), converted when it is written to the staged tile
template <int TYPE> struct RawPx { typedef float4 type; };
template <> struct RawPx<TC_U8> { typedef unsigned type; };
template <> struct RawPx<TC_U16> { typedef uint2 type; };
template <> struct RawPx<TC_F16> { typedef uint2 type; };
// streaming loads that do not allocate in L1: the staged pixels pass through once, the small per-axis tables stay cached
__device__ __forceinline__ float4 ld_stream(const float4* p)
{
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ uint2 ld_stream(const uint2* p)
{
    uint2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ unsigned ld_stream(const unsigned* p)
{
    unsigned v;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
template <int TYPE>
__device__ __forceinline__ typename RawPx<TYPE>::type load_raw(const void* __restrict__ base, size_t i)
{
    return ld_stream(reinterpret_cast<const typename RawPx<TYPE>::type*>(base) + i);
}
__device__ __forceinline__ float4 raw_px(float4 v, int) { return v; }
__device__ __forceinline__ float4 raw_px(unsigned v, int)
{
    return make_float4(u8_to_f(v & 255u), u8_to_f((v >> 8) & 255u), u8_to_f((v >> 16) & 255u), u8_to_f(v >> 24));
}
__device__ __forceinline__ float4 raw_px(uint2 v, int type)
{
    if (type == TC_U16) return make_float4(u16_to_f(v.x & 65535u), u16_to_f(v.x >> 16), u16_to_f(v.y & 65535u), u16_to_f(v.y >> 16));
    const float2 fa = __half22float2(*reinterpret_cast<__half2*>(&v.x)), fb = __half22float2(*reinterpret_cast<__half2*>(&v.y));
    return make_float4(fa.x, fa.y, fb.x, fb.y);
}

// The y tables of a block of BR destination rows, fetched with INDEPENDENT loads (one level of memory latency instead of
// first tap -> weight address -> weight): the rows' first taps and totals on lanes 0 .. BR - 1, the rows' BR x ytaps weights
// (contiguous in the table) spread over the lanes.  The expanded span x BR matrix is then scattered into shared memory.
template <int BR>
struct ColTab {
    int f;
    float tot;
    float w[BR];
};
template <int BR>
__device__ __forceinline__ void march_tab_load(ColTab<BR>& t, const int* __restrict__ yf, const float* __restrict__ yw,
                                               const float* __restrict__ yt, int ytaps, int yb, int yend)
{
    const int lane = threadIdx.x & 31;
    const int rows = min(BR, yend - yb);
    t.f = 0;
    t.tot = 0.0f;
    if (lane < rows) {
        t.f = yf[yb + lane];
        t.tot = yt[(yb + lane) * 2 + 0];
    }
    const float* w = yw + (size_t)yb * ytaps;
#pragma unroll
    for (int j = 0; j < BR; ++j) {
        const int i = lane + 32 * j;
        t.w[j] = i < rows * ytaps ? w[i] : 0.0f;
    }
}

template <int BR>
__device__ __forceinline__ void march_cols(ColTab<BR> t, const Img& dst, const float4* Mt, int ring, float4* wy, const int* __restrict__ yf,
                                           const float* __restrict__ yw, const float* __restrict__ yt, int ytaps, int ycur, int yend, int r0,
                                           int x0, int nx)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int yb = ycur + warp * BR; yb < yend; yb += RM_WARPS * BR) {
        const int rows = min(BR, yend - yb);
        const int fb = __shfl_sync(0xffffffffu, t.f, 0);
        const int span = min(__shfl_sync(0xffffffffu, t.f, rows - 1) + ytaps - fb, RM_SPAN);
        for (int k = lane; k < span; k += 32) wy[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncwarp();
#pragma unroll
        for (int j = 0; j < BR; ++j) {
            const int i = lane + 32 * j;
            int o = 0;
#pragma unroll
            for (int q = 1; q < BR; ++q) o += i >= q * ytaps ? 1 : 0;
            const int k = i - o * ytaps + __shfl_sync(0xffffffffu, t.f, o) - fb;
            const float tot = __shfl_sync(0xffffffffu, t.tot, o);
            if (i < rows * ytaps && tot != 0.0f && k < span) reinterpret_cast<float*>(wy + k)[o] = t.w[j];
        }
        __syncwarp();
        // the next block's tables travel while this block is filtered
        const int ybn = yb + RM_WARPS * BR;
        if (ybn < yend) march_tab_load<BR>(t, yf, yw, yt, ytaps, ybn, yend);
        const float4* col = Mt + lane * (ring + 1);
        const int s0 = (fb - r0) % ring;
        float4 acc[BR];
#pragma unroll
        for (int o = 0; o < BR; ++o) acc[o] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
        for (int k = 0; k < span; ++k) {
            const int slot = s0 + k;
            const float4 q = col[slot >= ring ? slot - ring : slot];
            const float4 wk = wy[k];
            acc[0] = fma4p(wk.x, q, acc[0]);
            acc[1] = fma4p(wk.y, q, acc[1]);
            if constexpr (BR > 2) {
                acc[2] = fma4p(wk.z, q, acc[2]);
                acc[3] = fma4p(wk.w, q, acc[3]);
            }
        }
        if (lane < nx) {
#pragma unroll
            for (int o = 0; o < BR; ++o)
                if (o < rows) store_px(dst.p, dst.type, (size_t)(yb + o) * dst.w + x0 + lane, acc[o]);
        }
        __syncwarp();
    }
}

// CS = 32-column spans of the staged tile (cols <= 32 CS).  PF: the loads of band n + 1 are issued into registers (in the
// storage type) right after band n has been written to the tile and land while band n runs through both passes -- two CTAs
// per SM (registers); without it three CTAs per SM where the tile leaves room, which hides the same latency with more warps.
template <int TYPE, int CS, bool PF>
__global__ void __launch_bounds__(RM_THREADS, PF ? 2 : 3) k_resize_march(const Img src_, const Img dst, const int* __restrict__ xf,
                                                                 const float* __restrict__ xw, const float* __restrict__ xt,
                                                                 const int* __restrict__ yf, const float* __restrict__ yw,
                                                                 const float* __restrict__ yt, const ResizeMarch T)
{
    extern __shared__ float4 rm_smem[];
    __shared__ int s_fxb[RM_XBLK], s_span[RM_XBLK], s_yend[RM_MAXBANDS], s_wsum[RM_WARPS];
    const Img src { src_.p, src_.w, src_.h, TYPE };
    const int swp = T.sw | 1;
    float4* S = rm_smem;                            // [RM_BAND][swp]
    float4* Mt = S + (size_t)RM_BAND * swp;         // [RM_W][ring + 1]: row-pass results, ring of source rows
    float4* Wx = Mt + RM_W * (T.ring + 1);          // [RM_XBLK][wspan]: expanded x weights of the strip's column blocks
    float4* Wy = Wx + RM_XBLK * T.wspan;            // [RM_WARPS][wspan]: a warp's expanded y weights of its current row block
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int x0 = blockIdx.x * RM_W, y0 = blockIdx.y * T.ch;
    const int nx = min(RM_W, dst.w - x0), y1 = min(y0 + T.ch, dst.h);
    const int fx0 = xf[x0];
    const int cols = min(min(xf[x0 + nx - 1] + T.xtaps - fx0, T.sw), 32 * CS);
    const int r0 = yf[y0];                          // first source row of the chunk (may lie above the frame: clamped when staged)
    const int rlast = yf[y1 - 1] + T.ytaps;         // one past the last source row the chunk needs

    constexpr int RPW = RM_BAND / RM_WARPS;         // rows per warp and band
    typename RawPx<TYPE>::type v[RPW][CS];
    int sx[CS];
#pragma unroll
    for (int q = 0; q < CS; ++q) sx[q] = min(max(fx0 + min(q * 32 + lane, cols - 1), 0), src.w - 1);
    auto issue = [&](int band0) {
#pragma unroll
        for (int rr = 0; rr < RPW; ++rr) {
            const size_t g = (size_t)min(max(band0 + warp + rr * RM_WARPS, 0), src.h - 1) * src.w;
#pragma unroll
            for (int q = 0; q < CS; ++q) v[rr][q] = load_raw<TYPE>(src.p, g + sx[q]);
        }
    };
    if (PF) issue(r0);

    // expanded x weights: Wx[b][k] = weights of staged column fxb[b] + k for the block's four outputs
    for (int i = tid; i < RM_XBLK * T.wspan; i += RM_THREADS) {
        const int b = i / T.wspan, k = i - b * T.wspan;
        float w[RM_B] = { 0.f, 0.f, 0.f, 0.f };
        if (b * RM_B < nx) {
            const int fb = xf[x0 + b * RM_B];
#pragma unroll
            for (int o = 0; o < RM_B; ++o) {
                const int x = x0 + b * RM_B + o;
                if (x < dst.w) {
                    const int t = k - (xf[x] - fb);
                    if (t >= 0 && t < T.xtaps && xt[x * 2 + 1] != 0.0f) w[o] = xw[(size_t)x * T.xtaps + t];
                }
            }
        }
        Wx[i] = make_float4(w[0], w[1], w[2], w[3]);
    }
    if (tid < RM_XBLK) {
        const int xa = min(x0 + tid * RM_B, dst.w - 1), xb = min(x0 + tid * RM_B + RM_B - 1, dst.w - 1);
        const int fb = min(xf[xa] - fx0, max(cols - 1, 0));
        s_fxb[tid] = fb;
        s_span[tid] = max(min(min(xf[xb] + T.xtaps - xf[xa], T.wspan), cols - fb), 0);
    }

    // destination rows complete after each band (yf is non-decreasing): counted once per CTA, so the walk below has no
    // dependent table look-ups in front of the column pass
    {
        const int nb = min((rlast - r0 + RM_BAND - 1) / RM_BAND, RM_MAXBANDS);
        s_yend[tid] = 0;
        __syncthreads();
        for (int y = y0 + tid; y < y1; y += RM_THREADS) atomicAdd(&s_yend[min((yf[y] + T.ytaps - r0 + RM_BAND - 1) / RM_BAND - 1, nb - 1)], 1);
        __syncthreads();
        int vsum = s_yend[tid];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, vsum, d);
            if (lane >= d) vsum += u;
        }
        if (lane == 31) s_wsum[warp] = vsum;
        __syncthreads();
        for (int q = 0; q < warp; ++q) vsum += s_wsum[q];
        s_yend[tid] = y0 + vsum;
    }

    int hi = r0; // source rows below `hi` have been through the row pass
    int ycur = y0;
    for (int band = 0; ycur < y1; ++band) {
        // ---- the band's pixels into the tile; then the next band's loads
        if (!PF) issue(hi);
#pragma unroll
        for (int rr = 0; rr < RPW; ++rr)
#pragma unroll
            for (int q = 0; q < CS; ++q) {
                const int c = q * 32 + lane;
                if (c < cols) S[(size_t)(warp + rr * RM_WARPS) * swp + c] = raw_px(v[rr][q], TYPE);
            }
        __syncthreads();
        if (PF && hi + RM_BAND < rlast) issue(hi + RM_BAND);
        const int yend = s_yend[min(band, RM_MAXBANDS - 1)];
        auto run = [&](auto brc) {
            constexpr int BR = decltype(brc)::value;
            // the tables of the warp's first block of destination rows travel during the row pass
            ColTab<BR> tab;
            tab.f = 0;
            tab.tot = 0.0f;
            if (ycur + warp * BR < yend) march_tab_load<BR>(tab, yf, yw, yt, T.ytaps, ycur + warp * BR, yend);
            // ---- row pass of the band into the ring
            if (warp * RM_B < nx) {
                const float4* row = S + (size_t)lane * swp + s_fxb[warp];
                const float4* w = Wx + warp * T.wspan;
                const int span = s_span[warp];
                float4 acc[RM_B];
#pragma unroll
                for (int o = 0; o < RM_B; ++o) acc[o] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
                for (int k = 0; k < span; ++k) {
                    const float4 q = row[k];
                    const float4 wk = w[k];
                    acc[0] = fma4p(wk.x, q, acc[0]);
                    acc[1] = fma4p(wk.y, q, acc[1]);
                    acc[2] = fma4p(wk.z, q, acc[2]);
                    acc[3] = fma4p(wk.w, q, acc[3]);
                }
                int slot = (hi - r0) % T.ring + lane;
                slot = slot >= T.ring ? slot - T.ring : slot;
#pragma unroll
                for (int o = 0; o < RM_B; ++o) Mt[(warp * RM_B + o) * (T.ring + 1) + slot] = acc[o];
            }
            __syncthreads();
            // ---- column pass over the destination rows whose window is now complete: a warp per block of BR rows
            march_cols<BR>(tab, dst, Mt, T.ring, Wy + warp * T.wspan, yf, yw, yt, T.ytaps, ycur, yend, r0, x0, nx);
        };
        // 2 rows per block when the band completes at most 16 rows, so that every warp has a block
        if (yend - ycur <= 2 * RM_WARPS) run(std::integral_constant<int, 2>());
        else run(std::integral_constant<int, RM_B>());
        hi += RM_BAND;
        ycur = yend;
    }
}

// Resize of float32 frames with warp-private rings (round 2, after k_gauss_march: the Blur march showed that a band staged by
// TMA, filtered by warps that never meet at a CTA barrier, beats the same work behind barriers).  A CTA owns NSUB sub-strips of
// 32 destination columns, each with 4 warps and its own staged tile (a TMA box is at most 128 pixels wide), and a chunk of rows;
// a warp owns 8 of the columns END TO END:
//   * bands of 16 source rows arrive by TMA (zero fill outside the frame; one staging buffer: the other CTA of the SM covers
//     the copy), counted on an mbarrier, handed back through a shared-memory counter;
//   * row pass: lane = (source row 0..15, half 0..1): the 4 outputs of one half of the warp's columns from their common span
//     (expanded span x 4 weights, one broadcast read per pixel read), into the warp's PRIVATE ring of 32 row slots;
//   * column pass: lane = (column 0..7, segment 0..3); a segment is a block of BRL consecutive destination rows filtered from
//     their common span of ring rows with an expanded span x BRL weight matrix -- built per band by the warp from the cached
//     tables, read as ONE broadcast per ring read (the eight lanes of a quarter warp share the segment);
//   * frame edges are index arithmetic: the row pass of an edge strip clamps its (uniform) column index, rows outside the frame
//     are never filtered, and a block whose windows reach over the top / bottom takes a tap-by-tap loop with a clamped row index.
// Same weights and ascending-tap accumulation as the kernels above: identical frames for finite pixels.
constexpr int WP_BAND = 16, WP_RING = 32, WP_RP = WP_RING + 1, WP_SPAN = 24, WP_COLS = 8;
// NSUB sub-strips of 32 destination columns (4 warps and one staged tile each) per CTA

struct ResizeWp {
    int sw;       // staged columns of one 32-column sub-strip (allocation); the staged row pitch is sw | 1 pixels
    int ch;       // destination rows per chunk
    int xtaps, ytaps;
    int stages;   // staging buffers (each two tiles): 2-4 with two CTAs per SM where they fit, else up to 4 with one CTA per SM
};

template <int BRL, int NSUB>
static size_t resize_wp_smem(int sw, int stages)
{
    const size_t tile = (((size_t)WP_BAND * (sw | 1) * 16) + 127) / 128 * 128;
    return (size_t)stages * NSUB * tile + (size_t)(4 * NSUB) * WP_COLS * WP_RP * 16 + (size_t)(8 * NSUB) * WP_SPAN * 16 + (size_t)(4 * NSUB) * 4 * WP_SPAN * BRL * 4 + 64;
}

template <int BRL, int NSUB>
__global__ void __launch_bounds__(128 * NSUB, NSUB == 1 ? 4 : 2) k_resize_wp(const Img src, const Img dst, const int* __restrict__ xf, const float* __restrict__ xw,
                                                              const float* __restrict__ xt, const int* __restrict__ yf,
                                                              const float* __restrict__ yw, const float* __restrict__ yt, const ResizeWp T,
                                                              const __grid_constant__ CUtensorMap map)
{
    extern __shared__ __align__(128) unsigned char wp_smem[];
    constexpr int WP_W = 32 * NSUB, WP_WARPS = 4 * NSUB, WP_THREADS = 128 * NSUB;
    __shared__ int s_fxb[WP_W / RM_B], s_span[WP_W / RM_B], s_yend[RM_MAXBANDS];
    const int swp = T.sw | 1;
    const size_t tile_bytes_al = (((size_t)WP_BAND * swp * 16) + 127) / 128 * 128;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const size_t stage_bytes = NSUB * tile_bytes_al;
    float4* rings = reinterpret_cast<float4*>(wp_smem + (size_t)T.stages * stage_bytes);
    float4* ring = rings + (size_t)warp * WP_COLS * WP_RP;                       // [8 columns][33]
    float4* Wx = rings + (size_t)WP_WARPS * WP_COLS * WP_RP;                     // [16 blocks][WP_SPAN]
    float* WyAll = reinterpret_cast<float*>(Wx + (WP_W / RM_B) * WP_SPAN);       // [8 warps][4 segments][WP_SPAN][BRL]
    float* Wy = WyAll + (size_t)warp * 4 * WP_SPAN * BRL;
    uint64_t* full = reinterpret_cast<uint64_t*>(WyAll + (size_t)WP_WARPS * 4 * WP_SPAN * BRL);
    int* released = reinterpret_cast<int*>(full + 4);

    const int x0 = blockIdx.x * WP_W, y0 = blockIdx.y * T.ch;
    const int nx = min(WP_W, dst.w - x0), y1 = min(y0 + T.ch, dst.h);
    // the two sub-strips: first staged column and staged width
    const int xa0 = x0, xa1 = min(x0 + 31, dst.w - 1);
    const int xb0 = min(x0 + 32, dst.w - 1), xb1 = min(x0 + 63, dst.w - 1);
    const int fxa = xf[xa0], fxbb = xf[xb0];
    const int cols_a = min(xf[xa1] + T.xtaps - fxa, T.sw), cols_b = min(xf[xb1] + T.xtaps - fxbb, T.sw);
    const bool has_b = NSUB > 1 && nx > 32;
    const int r0 = yf[y0];
    const int rlast = min(yf[y1 - 1] + T.ytaps, src.h);
    const int nb = min((rlast - r0 + WP_BAND - 1) / WP_BAND, RM_MAXBANDS);
    const unsigned tile_bytes = (unsigned)(WP_BAND * swp * 16);

    auto issue = [&](int band) {
        const int st = band % T.stages;
        unsigned char* to = wp_smem + (size_t)st * stage_bytes;
        mbar_expect_tx(&full[st], has_b ? 2 * tile_bytes : tile_bytes);
        tma_load_tile(to, &map, 2 * fxa, r0 + band * WP_BAND, &full[st]);
        if (has_b) tma_load_tile(to + tile_bytes_al, &map, 2 * fxbb, r0 + band * WP_BAND, &full[st]);
    };
    if (tid == 0) {
        for (int st = 0; st < T.stages; ++st) {
            mbar_init(&full[st], 1);
            released[st] = 0;
        }
        mbar_init_fence();
        for (int b = 0; b < T.stages && b < nb; ++b) issue(b);
    }
    // rings start as zeros: a block's expanded weights are zero outside an output's window, and zero times a stale slot must be zero
    for (int i = tid; i < WP_WARPS * WP_COLS * WP_RP; i += WP_THREADS) rings[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    // expanded x weights of the 16 blocks of 4 columns, relative to their sub-strip's tile
    for (int i = tid; i < (WP_W / RM_B) * WP_SPAN; i += WP_THREADS) {
        const int b = i / WP_SPAN, k = i - b * WP_SPAN;
        float w[RM_B] = { 0.f, 0.f, 0.f, 0.f };
        if (b * RM_B < nx) {
            const int fb = xf[x0 + b * RM_B];
#pragma unroll
            for (int o = 0; o < RM_B; ++o) {
                const int x = x0 + b * RM_B + o;
                if (x < dst.w) {
                    const int t = k - (xf[x] - fb);
                    if (t >= 0 && t < T.xtaps && xt[x * 2 + 1] != 0.0f) w[o] = xw[(size_t)x * T.xtaps + t];
                }
            }
        }
        Wx[i] = make_float4(w[0], w[1], w[2], w[3]);
    }
    if (tid < WP_W / RM_B) {
        const int xa = min(x0 + tid * RM_B, dst.w - 1), xb = min(x0 + tid * RM_B + RM_B - 1, dst.w - 1);
        const bool in_b = tid >= 8;
        const int base = in_b ? fxbb : fxa, cols = in_b ? cols_b : cols_a;
        const int fb = min(xf[xa] - base, max(cols - 1, 0));
        s_fxb[tid] = fb;
        s_span[tid] = max(min(min(xf[xb] + T.xtaps - xf[xa], WP_SPAN), cols - fb), 0);
    }
    // destination rows complete after each band of 16 source rows (see k_resize_march)
    for (int i = tid; i < RM_MAXBANDS; i += WP_THREADS) s_yend[i] = 0;
    __syncthreads();
    for (int y = y0 + tid; y < y1; y += WP_THREADS)
        atomicAdd(&s_yend[min(max((min(yf[y] + T.ytaps, src.h) - r0 + WP_BAND - 1) / WP_BAND - 1, 0), max(nb - 1, 0))], 1);
    __syncthreads();
    if (warp == 0) {
        // inclusive prefix over the RM_MAXBANDS counts: 8 consecutive entries per lane, then a warp scan of the lane sums
        constexpr int PER = RM_MAXBANDS / 32;
        int v[PER], sum = 0;
#pragma unroll
        for (int q = 0; q < PER; ++q) {
            sum += s_yend[lane * PER + q];
            v[q] = sum;
        }
        int incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += u;
        }
        const int before = incl - sum;
#pragma unroll
        for (int q = 0; q < PER; ++q) s_yend[lane * PER + q] = y0 + before + v[q];
    }
    __syncthreads();

    // this lane's roles
    const int rrow = lane & 15, half = lane >> 4;          // row pass
    const int blk = warp * 2 + half;                        // block of 4 destination columns
    const bool in_b = warp >= 4;
    const size_t tile_off = in_b ? tile_bytes_al : 0;
    const int tile_fx = in_b ? fxbb : fxa, tile_cols = in_b ? cols_b : cols_a;
    const bool xedge = tile_fx < 0 || tile_fx + tile_cols > src.w;
    const bool live = blk * RM_B < nx;
    const int fxb = s_fxb[blk], span = s_span[blk];
    const float4* wxb = Wx + blk * WP_SPAN;
    const int ccol = lane & 7, seg = lane >> 3;             // column pass
    const int xo = x0 + warp * WP_COLS + ccol;
    const float4* ringcol = ring + ccol * WP_RP;

    int ycur = y0;
    for (int band = 0; band < nb; ++band) {
        const int hi = r0 + band * WP_BAND;
        const int yend = s_yend[band];
        const int st = band % T.stages;
        const unsigned char* tile = wp_smem + (size_t)st * stage_bytes + tile_off;
        mbar_wait(&full[st], (band / T.stages) & 1);
        // ---- row pass
        if (live && hi + rrow < src.h && hi + rrow >= 0) {
            const float4* row = reinterpret_cast<const float4*>(tile) + (size_t)rrow * swp;
            float4 acc[RM_B];
#pragma unroll
            for (int q = 0; q < RM_B; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (!xedge) {
                const float4* p = row + fxb;
#pragma unroll 4
                for (int k = 0; k < span; ++k) {
                    const float4 q = p[k];
                    const float4 wk = wxb[k];
                    acc[0] = fma4p(wk.x, q, acc[0]);
                    acc[1] = fma4p(wk.y, q, acc[1]);
                    acc[2] = fma4p(wk.z, q, acc[2]);
                    acc[3] = fma4p(wk.w, q, acc[3]);
                }
            } else {
#pragma unroll 2
                for (int k = 0; k < span; ++k) {
                    const float4 q = row[min(max(tile_fx + fxb + k, 0), src.w - 1) - tile_fx];
                    const float4 wk = wxb[k];
                    acc[0] = fma4p(wk.x, q, acc[0]);
                    acc[1] = fma4p(wk.y, q, acc[1]);
                    acc[2] = fma4p(wk.z, q, acc[2]);
                    acc[3] = fma4p(wk.w, q, acc[3]);
                }
            }
            const int slot = (hi - r0 + rrow) & (WP_RING - 1);
#pragma unroll
            for (int q = 0; q < RM_B; ++q) ring[(half * RM_B + q) * WP_RP + slot] = acc[q];
        }
        __syncwarp();
        // ---- hand the tile back; the last warp to do so starts the next band's copy
        if (lane == 0 && release_count(&released[st]) == WP_WARPS - 1) {
            released[st] = 0;
            if (band + T.stages < nb) {
                fence_proxy_async();
                issue(band + T.stages);
            }
        }
        // ---- column pass over the rows this band completed, 4 segments of BRL rows per pass
        for (int base = ycur; base < yend; base += 4 * BRL) {
            const int nrows = min(4 * BRL, yend - base);
            // first taps and totals of the pass's rows on lanes 0 .. nrows - 1
            int fy = 0;
            float tot = 0.0f;
            if (lane < nrows) {
                fy = yf[base + lane];
                tot = yt[(base + lane) * 2];
            }
            const int seg_first = min(seg * BRL, nrows - 1), seg_last = min(seg * BRL + BRL - 1, nrows - 1);
            const int fb = __shfl_sync(0xffffffffu, fy, seg_first);
            const int fl = __shfl_sync(0xffffffffu, fy, seg_last);
            const bool seg_live = seg * BRL < nrows;
            const int sspan = seg_live ? min(fl + T.ytaps - fb, WP_SPAN) : 0;
            int mspan = sspan;
#pragma unroll
            for (int d = 8; d < 32; d <<= 1) mspan = max(mspan, __shfl_xor_sync(0xffffffffu, mspan, d));
            // windows that reach over the top / bottom of the frame: tap by tap with a clamped row index
            const int f0 = __shfl_sync(0xffffffffu, fy, 0), fe = __shfl_sync(0xffffffffu, fy, nrows - 1);
            if (f0 < 0 || fe + T.ytaps > src.h) {
                for (int o = 0; o < BRL; ++o) {
                    const int r = seg * BRL + o;
                    if (r >= nrows) break;
                    const int y = base + r;
                    const int fyr = yf[y];
                    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (yt[y * 2] != 0.0f)
                        for (int t = 0; t < T.ytaps; ++t)
                            acc = fma4p(yw[(size_t)y * T.ytaps + t], ringcol[(min(max(fyr + t, 0), src.h - 1) - r0) & (WP_RING - 1)], acc);
                    if (xo < dst.w && warp * WP_COLS + ccol < nx) store_px(dst.p, dst.type, (size_t)y * dst.w + xo, acc);
                }
                __syncwarp();
                continue;
            }
            // expanded weights: zero, then scatter the rows' taps (contiguous in the table) to their place in their segment
            for (int i = lane; i < 4 * WP_SPAN * BRL; i += 32) Wy[i] = 0.0f;
            __syncwarp();
            const float* wsrc = yw + (size_t)base * T.ytaps;
            // (shuffles under a divergent loop bound are avoided: every lane walks the same number of iterations below)
            const int total = nrows * T.ytaps;
            for (int i0 = 0; i0 < total; i0 += 32) {
                const int i = min(i0 + lane, total - 1);
                const int r = i / T.ytaps, t = i - r * T.ytaps;
                const int sg = r / BRL, o = r - sg * BRL;
                const int fr = __shfl_sync(0xffffffffu, fy, r);
                const int fs = __shfl_sync(0xffffffffu, fy, sg * BRL);
                const float tr = __shfl_sync(0xffffffffu, tot, r);
                const int k = t + fr - fs;
                if (i0 + lane < total && tr != 0.0f && k < WP_SPAN) Wy[(sg * WP_SPAN + k) * BRL + o] = wsrc[i];
            }
            __syncwarp();
            // the block of BRL rows of this lane's column
            float4 acc[BRL];
#pragma unroll
            for (int o = 0; o < BRL; ++o) acc[o] = make_float4(0.f, 0.f, 0.f, 0.f);
            const float* wseg = Wy + (size_t)seg * WP_SPAN * BRL;
            const int s0 = (fb - r0) & (WP_RING - 1);
#pragma unroll 2
            for (int k = 0; k < mspan; ++k) {
                const float4 q = ringcol[(s0 + k) & (WP_RING - 1)];
                if constexpr (BRL == 2) {
                    const float2 wk = *reinterpret_cast<const float2*>(wseg + k * 2);
                    acc[0] = fma4p(wk.x, q, acc[0]);
                    acc[1] = fma4p(wk.y, q, acc[1]);
                } else {
#pragma unroll
                    for (int v = 0; v < BRL / 4; ++v) {
                        const float4 wk = *reinterpret_cast<const float4*>(wseg + k * BRL + v * 4);
                        acc[v * 4 + 0] = fma4p(wk.x, q, acc[v * 4 + 0]);
                        acc[v * 4 + 1] = fma4p(wk.y, q, acc[v * 4 + 1]);
                        acc[v * 4 + 2] = fma4p(wk.z, q, acc[v * 4 + 2]);
                        acc[v * 4 + 3] = fma4p(wk.w, q, acc[v * 4 + 3]);
                    }
                }
            }
            if (warp * WP_COLS + ccol < nx) {
#pragma unroll
                for (int o = 0; o < BRL; ++o) {
                    const int r = seg * BRL + o;
                    if (r < nrows) store_px(dst.p, dst.type, (size_t)(base + r) * dst.w + xo, acc[o]);
                }
            }
            __syncwarp();
        }
        ycur = yend;
        __syncwarp();
    }
}

struct RotParams {
    float m[9];  // inverse matrix, row-vector convention
    int id;      // filter id
    float fw;    // filter width
    float fscale;
};

constexpr int ROT_MAX_TAPS = 16;

// MAXT = taps per axis the footprint can span: 8 for the default lanczos3 (width 6), 16 otherwise.
// FILT selects the filter at compile time: with both filters inlined at every tap the kernel body
// outgrew the instruction cache (45 % of the stall samples were instruction fetches).
template <int FILT>
__device__ __forceinline__ float rot_filt(float scale, float x)
{
    return FILT == 0 ? lanczos3(__fmul_rn(x, scale)) : bh1d(__fmul_rn(x, scale));
}

template <int MAXT, int FILT, int TYPE>
__global__ void __launch_bounds__(256) k_rotate(const Img src_, const Img dst, const __grid_constant__ RotParams P)
{
    constexpr int ROT_MAX_TAPS = MAXT;
    const Img src { src_.p, src_.w, src_.h, TYPE }; // storage type known at compile time: one load path per tap
    // a warp covers an 8 x 4 pixel block, not a 32 x 1 row: under rotation a row of the output is a
    // diagonal of the source (32 cache lines per tap), a compact block maps to a compact footprint
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int x = blockIdx.x * 32 + (warp & 3) * 8 + (lane & 7);
    const int y = blockIdx.y * 8 + (warp >> 2) * 4 + (lane >> 3);
    if (x >= dst.w || y >= dst.h) return;
    const float px = __fadd_rn((float)x, 0.5f), py = __fadd_rn((float)y, 0.5f);
    const float s = __fadd_rn(__fadd_rn(__fmul_rn(px, P.m[0]), __fmul_rn(py, P.m[3])), P.m[6]);
    const float t = __fadd_rn(__fadd_rn(__fmul_rn(px, P.m[1]), __fmul_rn(py, P.m[4])), P.m[7]);
    const float ds = fmaxf(1.0f, fmaxf(fabsf(P.m[0]), fabsf(P.m[3])));
    const float dt = fmaxf(1.0f, fmaxf(fabsf(P.m[1]), fabsf(P.m[4])));
    const float ds_inv = __fdiv_rn(1.0f, ds), dt_inv = __fdiv_rn(1.0f, dt);
    const float rs = __fmul_rn(__fmul_rn(0.5f, ds), P.fw), rt = __fmul_rn(__fmul_rn(0.5f, dt), P.fw);
    const int smin = (int)floorf(__fsub_rn(s, rs)), smax = (int)ceilf(__fadd_rn(s, rs));
    const int tmin = (int)floorf(__fsub_rn(t, rt)), tmax = (int)ceilf(__fadd_rn(t, rt));
    const int ns = min(smax - smin, ROT_MAX_TAPS), nt = min(tmax - tmin, ROT_MAX_TAPS);
    float wx[ROT_MAX_TAPS];
#pragma unroll
    for (int i = 0; i < ROT_MAX_TAPS; ++i)
        wx[i] = i < ns ? rot_filt<FILT>(P.fscale, __fmul_rn(ds_inv, __fsub_rn(__fadd_rn((float)(smin + i), 0.5f), s))) : 0.0f;
    float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
    float total = 0.0f;
    // The weights follow OIIO's unfused float sequence; the accumulation uses fused multiply-adds
    // (tolerance op: <= 1e-5).  Footprints fully inside the source skip the per-tap bounds test.
    const bool inside = smin >= 0 && tmin >= 0 && smin + ns <= src.w && tmin + nt <= src.h;
#pragma unroll 1
    for (int j = 0; j < nt; ++j) {
        const float wy = rot_filt<FILT>(P.fscale, __fmul_rn(dt_inv, __fsub_rn(__fadd_rn((float)(tmin + j), 0.5f), t)));
        const size_t row = (size_t)(tmin + j) * src.w + smin;
#pragma unroll
        for (int i = 0; i < ROT_MAX_TAPS; ++i) {
            if (i < ns) {
                const float w = __fmul_rn(wx[i], wy);
                const float4 p = inside ? load_px(src.p, src.type, row + i) : load_black(src, smin + i, tmin + j);
                sum.x = fmaf(w, p.x, sum.x);
                sum.y = fmaf(w, p.y, sum.y);
                sum.z = fmaf(w, p.z, sum.z);
                sum.w = fmaf(w, p.w, sum.w);
                total = __fadd_rn(total, w);
            }
        }
    }
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (total > 0.0f)
        o = make_float4(__fdiv_rn(sum.x, total), __fdiv_rn(sum.y, total), __fdiv_rn(sum.z, total), __fdiv_rn(sum.w, total));
    store_px(dst.p, dst.type, (size_t)y * dst.w + x, o);
}

// Filter evaluation for the tile kernel (tolerance path): OIIO's parabola sinpi with its multiply-adds
// fused, the quotient 3 / (pi x)^2 through the approximate reciprocal instead of an IEEE division with its
// slow-path call, selects instead of branches -- 25 instructions instead of 50.  The weights differ from
// OIIO's by a few 1e-7 relative, the normalised result by less than 1e-6.
__device__ __forceinline__ float fast_sinpi_fused(float x)
{
    const float z = __fsub_rn(x, __fsub_rn(__fadd_rn(x, 25165824.0f), 25165824.0f));
    const float y = fmaf(-z, fabsf(z), z);
    return __fmul_rn(y, fmaf(3.584135056f, fabsf(y), 3.10396624f));
}
template <int FILT>
__device__ __forceinline__ float rot_filt_fast(float x)
{
    if (FILT != 0) return bh1d(x);
    x = fabsf(x);
    const float inv = __fmul_rn(0.30396355092701331f, rcp_approx(__fmul_rn(x, x))); // 3 / (pi x)^2
    float w = __fmul_rn(__fmul_rn(inv, fast_sinpi_fused(x)), fast_sinpi_fused(__fmul_rn(x, 1.0f / 3.0f)));
    w = x < 0.0001f ? 1.0f : w;
    return x > 3.0f ? 0.0f : w;
}
// true where the filter is exactly zero (lanczos3: sinpi(3) is exactly 0 in the parabola form)
template <int FILT>
__device__ __forceinline__ bool rot_filt_zero(float x)
{
    return FILT == 0 ? fabsf(x) >= 3.0f : (x < -1.0f || x > 1.0f);
}

// The six live lanczos3 weights of one axis at x_k = b0 + k, k = 0 .. 5, for unit tap spacing and b0 in (-3, -2].  OIIO's
// parabola sinpi is odd with period 2, so |fast_sinpi(x_k)| is the same for all six taps and its sign runs + - + + - +:
// ONE evaluation instead of six; |x_k| / 3 <= 1 needs no range reduction.  ~11 instructions per weight instead of ~25,
// each within float rounding (a few 1e-7 relative) of rot_filt_fast.
__device__ __forceinline__ void lanczos3_six(float b0, float (&w)[6])
{
    const float A = __fmul_rn(fabsf(fast_sinpi_fused(b0)), 0.30396355092701331f); // |sinpi| * 3 / pi^2
    // sinpi(x / 3): |x_k| / 3 and |x_(k+3)| / 3 add up to 1 and the parabola u (1 - u) is symmetric about 1/2 -- three evaluations
    float f3s[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float u = __fmul_rn(fabsf(__fadd_rn(b0, (float)k)), 1.0f / 3.0f);
        const float y = fmaf(-u, u, u);
        f3s[k] = __fmul_rn(y, fmaf(3.584135056f, y, 3.10396624f));
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const float x = fabsf(__fadd_rn(b0, (float)k));
        const float f3 = f3s[k < 3 ? k : k - 3];
        float v = __fmul_rn(__fmul_rn(rcp_approx(__fmul_rn(x, x)), (k == 1 || k == 4) ? -A : A), f3);
        if (k == 2 || k == 3) v = x < 0.0001f ? 1.0f : v;
        w[k] = v;
    }
}

// The same gather through a shared-memory tile (the north-star layout for Rotate).  A CTA owns a 16 x 16 block
// of the destination; the inverse map is affine, so the block's source footprint is the bounding box of its
// four corner footprints.  The box is staged once with black outside the source -- which also removes the
// per-tap bounds test -- and every pixel then filters out of shared memory: 4 source pixels loaded per
// destination pixel instead of 49-64 L1 gathers.  Weights to a few 1e-7 of k_rotate's (rot_filt_fast);
// live taps only, rows reduced first (see the kernel).
constexpr int RTT = 16; // destination block edge

struct RotTile {
    int cols, rows, pitch;                       // allocation of the shared source tile; row pitch in pixels
    float smn_off, smx_off, tmn_off, tmx_off;    // extent of a block's four corner positions relative to its first pixel's
    int generic;                                 // k_rotate_quad: every quad through the per-pixel form (TOUCAN_ROTATE_QUAD=generic, tests)
    int raw_off;                                 // k_rotate_quad: byte offset of the storage-type tile the copy engine fills
};

// TMA: the source box is fetched by ONE tensor copy (float32 sources; zero fill outside the frame IS rotate's "black outside
// the source") issued by thread 0 and awaited on an mbarrier only after every thread has evaluated its 12 filter weights --
// the copy's latency hides behind them.  Otherwise a warp per source row stages the box with bounds-checked loads.
template <int MAXT, int FILT, int TYPE, bool TMA>
__global__ void __launch_bounds__(RTT * RTT) k_rotate_tile(const Img src_, const Img dst, const __grid_constant__ RotParams P, const RotTile T,
                                                           const __grid_constant__ CUtensorMap map)
{
    extern __shared__ __align__(128) float4 rot_smem[];
    __shared__ uint64_t bar;
    const Img src { src_.p, src_.w, src_.h, TYPE };
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * RTT, y0 = blockIdx.y * RTT;
    const float ds = fmaxf(1.0f, fmaxf(fabsf(P.m[0]), fabsf(P.m[3])));
    const float dt = fmaxf(1.0f, fmaxf(fabsf(P.m[1]), fabsf(P.m[4])));
    const float rs = __fmul_rn(__fmul_rn(0.5f, ds), P.fw), rt = __fmul_rn(__fmul_rn(0.5f, dt), P.fw);
    // footprint of the block: the affine map takes its extremes at the corner pixels -- the first pixel's position plus
    // the host's corner offsets (two pixels of slack cover the rounding of either form)
    const float px0 = __fadd_rn((float)x0, 0.5f), py0 = __fadd_rn((float)y0, 0.5f);
    const float s00 = __fadd_rn(__fadd_rn(__fmul_rn(px0, P.m[0]), __fmul_rn(py0, P.m[3])), P.m[6]);
    const float t00 = __fadd_rn(__fadd_rn(__fmul_rn(px0, P.m[1]), __fmul_rn(py0, P.m[4])), P.m[7]);
    const int s0 = (int)floorf(s00 + T.smn_off - rs) - 1, t0 = (int)floorf(t00 + T.tmn_off - rt) - 1;
    const int cols = min((int)ceilf(s00 + T.smx_off + rs) + 1 - s0, T.cols), rows = min((int)ceilf(t00 + T.tmx_off + rt) + 1 - t0, T.rows);
    const int pitch = T.pitch;
    if (TMA) {
        if (tid == 0) {
            mbar_init(&bar, 1);
            mbar_init_fence();
            mbar_expect_tx(&bar, (unsigned)(T.rows * T.pitch * 16));
            tma_load_tile(rot_smem, &map, 2 * s0, t0, &bar);
        }
    } else {
        // stage: a warp per source row, black outside the source
        for (int r = tid >> 5; r < rows; r += RTT * RTT / 32)
            for (int c = tid & 31; c < cols; c += 32) rot_smem[r * pitch + c] = load_black(src, s0 + c, t0 + r);
    }
    __syncthreads();

    // a warp covers 8 x 4 destination pixels and each quarter warp (one 128-bit shared-memory phase) a compact 4 x 2 block:
    // its eight footprints start inside a ~3 x 3 patch of the tile instead of along a diagonal of 6 (fewer bank conflicts)
    const int lane = tid & 31, warp = tid >> 5;
    const int x = x0 + (warp & 1) * 8 + ((lane >> 3) & 1) * 4 + (lane & 3), y = y0 + (warp >> 1) * 4 + (lane >> 4) * 2 + ((lane >> 2) & 1);
    if (x >= dst.w || y >= dst.h) return;
    const float px = __fadd_rn((float)x, 0.5f), py = __fadd_rn((float)y, 0.5f);
    const float s = __fadd_rn(__fadd_rn(__fmul_rn(px, P.m[0]), __fmul_rn(py, P.m[3])), P.m[6]);
    const float t = __fadd_rn(__fadd_rn(__fmul_rn(px, P.m[1]), __fmul_rn(py, P.m[4])), P.m[7]);
    const float ds_inv = __fdiv_rn(1.0f, ds), dt_inv = __fdiv_rn(1.0f, dt);
    const int smin = (int)floorf(__fsub_rn(s, rs)), smax = (int)ceilf(__fadd_rn(s, rs));
    const int tmin = (int)floorf(__fsub_rn(t, rt)), tmax = (int)ceilf(__fadd_rn(t, rt));
    const int ns = min(smax - smin, MAXT), nt = min(tmax - tmin, MAXT);
    // indices stay inside the staged box by construction; the clamps only guard the allocation
    const int cbase = min(max(smin - s0, 0), max(cols - ns, 0));
    const int rbase = min(max(tmin - t0, 0), max(rows - nt, 0));
    // filter argument of tap i: ((smin + i + 0.5) - s) * ds_inv * fscale; the difference is exact in float
    const float bx = __fsub_rn(__fadd_rn((float)smin, 0.5f), s), by = __fsub_rn(__fadd_rn((float)tmin, 0.5f), t);
    const float kx = __fmul_rn(ds_inv, P.fscale), ky = __fmul_rn(dt_inv, P.fscale);
    // A lanczos3 footprint of width 6 spans 7 pixel centres, but the filter is zero at distance >= 3, so at
    // most 6 of the 7 weights per axis are live and the dead one sits at either end (which end depends on
    // the fractional position, i.e. on the thread).  Starting the window one tap later where the leading
    // weight is zero gives every thread the same 6 x 6 loop: 12 filter evaluations instead of 15, no per-tap
    // predicate, no divergence, 36 taps instead of 49.  Rows are reduced with the x weights first and folded
    // in with one multiply-add per row (2 issue slots per tap); the weight total factors the same way.
    // Dropped taps have weight exactly 0, so the sums differ from the tap-by-tap order only by float
    // rounding (tolerance path, <= 1e-5).
    constexpr int LIVE = 6;
    const int i0 = rot_filt_zero<FILT>(__fmul_rn(bx, kx)) ? 1 : 0, j0 = rot_filt_zero<FILT>(__fmul_rn(by, ky)) ? 1 : 0;
    const bool fast = ns <= LIVE + 1 && nt <= LIVE + 1 && cbase + i0 + LIVE <= cols && rbase + j0 + LIVE <= rows &&
                      (i0 || ns <= LIVE || rot_filt_zero<FILT>(__fmul_rn(__fadd_rn(bx, (float)LIVE), kx))) &&
                      (j0 || nt <= LIVE || rot_filt_zero<FILT>(__fmul_rn(__fadd_rn(by, (float)LIVE), ky)));
    float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
    float total = 0.0f;
    const float bx0 = __fadd_rn(bx, (float)i0), by0 = __fadd_rn(by, (float)j0);
    // lanczos3 at unit tap spacing (every plain rotation): the shared-sinpi evaluation; anything else of lanczos3 goes tap by tap
    const bool six = FILT != 0 || (kx == 1.0f && ky == 1.0f && bx0 > -3.0f && bx0 <= -2.0f && by0 > -3.0f && by0 <= -2.0f);
    if (fast && six) {
        float ax[LIVE], ay[LIVE];
        float sx = 0.0f, sy = 0.0f;
        if (FILT == 0) {
            lanczos3_six(bx0, ax);
            lanczos3_six(by0, ay);
        }
#pragma unroll
        for (int k = 0; k < LIVE; ++k) {
            if (FILT != 0) {
                ax[k] = rot_filt_fast<FILT>(__fmul_rn(__fadd_rn(bx0, (float)k), kx));
                ay[k] = rot_filt_fast<FILT>(__fmul_rn(__fadd_rn(by0, (float)k), ky));
            }
            ax[k] = i0 + k < ns ? ax[k] : 0.0f;
            ay[k] = j0 + k < nt ? ay[k] : 0.0f;
            sx = __fadd_rn(sx, ax[k]);
            sy = __fadd_rn(sy, ay[k]);
        }
        total = __fmul_rn(sx, sy);
        if (TMA) mbar_wait(&bar, 0);
        const float4* row = rot_smem + (rbase + j0) * pitch + cbase + i0;
#pragma unroll
        for (int j = 0; j < LIVE; ++j, row += pitch) {
            float4 r = mul4p(ax[0], row[0]);
#pragma unroll
            for (int i = 1; i < LIVE; ++i) r = fma4p(ax[i], row[i], r);
            sum = j == 0 ? mul4p(ay[0], r) : fma4p(ay[j], r, sum);
        }
    } else {
        // any other footprint (both end taps live, wider filters): tap by tap
        float wx[MAXT];
#pragma unroll
        for (int i = 0; i < MAXT; ++i) wx[i] = i < ns ? rot_filt_fast<FILT>(__fmul_rn(__fadd_rn(bx, (float)i), kx)) : 0.0f;
        if (TMA) mbar_wait(&bar, 0);
#pragma unroll 1
        for (int j = 0; j < nt; ++j) {
            const float wy = rot_filt_fast<FILT>(__fmul_rn(__fadd_rn(by, (float)j), ky));
            const float4* row = rot_smem + (rbase + j) * pitch + cbase;
#pragma unroll
            for (int i = 0; i < MAXT; ++i) {
                if (i < ns) {
                    const float w = __fmul_rn(wx[i], wy);
                    sum = fma4p(w, row[i], sum);
                    total = __fadd_rn(total, w);
                }
            }
        }
    }
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (total > 0.0f) o = mul4p(rcp_approx(total), sum);
    store_px(dst.p, dst.type, (size_t)y * dst.w + x, o);
}

// Four destination pixels per thread (round 2): a thread owns a 2 x 2 block of the destination.  Under a rotation its four
// footprints start within two source pixels of each other on either axis, so one 8 x 8 window of the staged tile holds all four
// 6 x 6 live windows: 16 shared-memory pixel reads per destination pixel instead of 36.  Each pixel's six weights per axis are
// placed in an 8-wide row at the pixel's offset in the shared window (zeros elsewhere), the window rows are reduced with the four
// x rows and folded in with the y weights as in k_rotate_tile.  lanczos3 at its default width and unit tap spacing, the tile
// through the TMA copy in the storage type (8 / 16-bit and half tiles converted once per staged pixel); anything else -- and
// any quad whose pixels fall outside those conditions by float rounding -- takes the per-pixel forms (rot_px_generic here,
// k_rotate_tile / k_rotate otherwise).  Same weights as k_rotate_tile.
constexpr int RQ_BW = 32, RQ_BH = 16, RQ_THREADS = 128, RQ_WIN = 8;

__device__ __noinline__ void rot_px_generic(const float4* tile, int pitch, int rows, int s0, int t0, float s, float t, const Img dst, int x, int y)
{
    const int smin = (int)floorf(__fsub_rn(s, 3.0f)), smax = (int)ceilf(__fadd_rn(s, 3.0f));
    const int tmin = (int)floorf(__fsub_rn(t, 3.0f)), tmax = (int)ceilf(__fadd_rn(t, 3.0f));
    const int ns = min(smax - smin, 8), nt = min(tmax - tmin, 8);
    const int cbase = min(max(smin - s0, 0), max(pitch - ns, 0)), rbase = min(max(tmin - t0, 0), max(rows - nt, 0));
    const float bx = __fsub_rn(__fadd_rn((float)smin, 0.5f), s), by = __fsub_rn(__fadd_rn((float)tmin, 0.5f), t);
    float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
    float total = 0.0f;
#pragma unroll 1
    for (int j = 0; j < nt; ++j) {
        const float wy = rot_filt_fast<0>(__fadd_rn(by, (float)j));
        const float4* row = tile + (rbase + j) * pitch + cbase;
#pragma unroll 1
        for (int i = 0; i < ns; ++i) {
            const float w = __fmul_rn(rot_filt_fast<0>(__fadd_rn(bx, (float)i)), wy);
            sum = fma4p(w, row[i], sum);
            total = __fadd_rn(total, w);
        }
    }
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (total > 0.0f) o = mul4p(rcp_approx(total), sum);
    store_px(dst.p, dst.type, (size_t)y * dst.w + x, o);
}

template <int TYPE>
__global__ void __launch_bounds__(RQ_THREADS, 4) k_rotate_quad(const Img dst_, const __grid_constant__ RotParams P, const RotTile T,
                                                                const __grid_constant__ CUtensorMap map)
{
    extern __shared__ __align__(128) float4 rot_smem[];
    __shared__ uint64_t bar;
    const Img dst { dst_.p, dst_.w, dst_.h, TYPE };
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * RQ_BW, y0 = blockIdx.y * RQ_BH;
    const float px0 = __fadd_rn((float)x0, 0.5f), py0 = __fadd_rn((float)y0, 0.5f);
    const float s00 = __fadd_rn(__fadd_rn(__fmul_rn(px0, P.m[0]), __fmul_rn(py0, P.m[3])), P.m[6]);
    const float t00 = __fadd_rn(__fadd_rn(__fmul_rn(px0, P.m[1]), __fmul_rn(py0, P.m[4])), P.m[7]);
    // The source box arrives by ONE TMA tensor copy in the storage type (zero fill outside the frame IS rotate's black
    // surround).  A box row starts on a 16-byte boundary of the source row: the first column is rounded down to 4 pixels of
    // 8 bits, 2 pixels of 16 bits; the host allocates the extra columns.
    typedef typename RawPx<TYPE>::type Raw;
    constexpr int ALIGN = 16 / (int)sizeof(Raw);
    const int s0 = ((int)floorf(s00 + T.smn_off - 3.0f) - 1) & ~(ALIGN > 1 ? ALIGN - 1 : 0), t0 = (int)floorf(t00 + T.tmn_off - 3.0f) - 1;
    const int pitch = T.pitch;
    Raw* raw = reinterpret_cast<Raw*>(reinterpret_cast<unsigned char*>(rot_smem) + T.raw_off);
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_init_fence();
        mbar_expect_tx(&bar, (unsigned)(T.rows * T.pitch * sizeof(Raw)));
        tma_load_tile(raw, &map, TYPE == TC_F32 ? 2 * s0 : s0, t0, &bar);
    }
    __syncthreads();

    // warps 2 x 2 over the block (16 x 8 pixels each), quarter warps 2 x 2 over the warp, threads 4 x 2 over the quarter
    const int lane = tid & 31, warp = tid >> 5, q = lane >> 3, l = lane & 7;
    constexpr int WX = RQ_BW / 16;
    const int x = x0 + (warp % WX) * 16 + (q & 1) * 8 + (l & 3) * 2, y = y0 + (warp / WX) * 8 + (q >> 1) * 4 + (l >> 2) * 2;
    const bool live = x < dst.w && y < dst.h;
    if (TYPE == TC_F32 && !live) return; // (8 / 16-bit tiles are converted by the whole CTA below)

    float se[4], te[4];
    int cs[4], rs[4];
    bool ok = true;
    float exw[4][RQ_WIN], eyw[4][RQ_WIN], total[4];
    int cmin = 1 << 30, rmin = 1 << 30;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const float px = __fadd_rn((float)(x + (p & 1)), 0.5f), py = __fadd_rn((float)(y + (p >> 1)), 0.5f);
        const float s = __fadd_rn(__fadd_rn(__fmul_rn(px, P.m[0]), __fmul_rn(py, P.m[3])), P.m[6]);
        const float t = __fadd_rn(__fadd_rn(__fmul_rn(px, P.m[1]), __fmul_rn(py, P.m[4])), P.m[7]);
        se[p] = s;
        te[p] = t;
        const int smin = (int)floorf(__fsub_rn(s, 3.0f)), smax = (int)ceilf(__fadd_rn(s, 3.0f));
        const int tmin = (int)floorf(__fsub_rn(t, 3.0f)), tmax = (int)ceilf(__fadd_rn(t, 3.0f));
        const int ns = smax - smin, nt = tmax - tmin;
        const float bx = __fsub_rn(__fadd_rn((float)smin, 0.5f), s), by = __fsub_rn(__fadd_rn((float)tmin, 0.5f), t);
        const int i0 = fabsf(bx) >= 3.0f ? 1 : 0, j0 = fabsf(by) >= 3.0f ? 1 : 0;
        const float bx0 = __fadd_rn(bx, (float)i0), by0 = __fadd_rn(by, (float)j0);
        // the conditions of k_rotate_tile's six-weight form, all six taps inside the footprint (anything else is a rounding
        // accident of the position: per-pixel form)
        ok = ok & (ns <= 7) & (nt <= 7) & (i0 + 6 <= ns) & (j0 + 6 <= nt) & ((i0 != 0) | (ns <= 6) | (fabsf(__fadd_rn(bx, 6.0f)) >= 3.0f)) &
             ((j0 != 0) | (nt <= 6) | (fabsf(__fadd_rn(by, 6.0f)) >= 3.0f)) & (bx0 > -3.0f) & (bx0 <= -2.0f) & (by0 > -3.0f) & (by0 <= -2.0f);
        cs[p] = smin + i0;
        rs[p] = tmin + j0;
        cmin = min(cmin, cs[p]);
        rmin = min(rmin, rs[p]);
        float ax[6], ay[6];
        lanczos3_six(bx0, ax);
        lanczos3_six(by0, ay);
        float sx = 0.0f, sy = 0.0f;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            sx = __fadd_rn(sx, ax[k]);
            sy = __fadd_rn(sy, ay[k]);
            exw[p][k] = ax[k];
            eyw[p][k] = ay[k];
        }
        total[p] = __fmul_rn(sx, sy);
        exw[p][6] = exw[p][7] = eyw[p][6] = eyw[p][7] = 0.0f;
    }
    const int cb = cmin - s0, rb = rmin - t0;
    ok = ok & (cb >= 0) & (cb + RQ_WIN <= pitch) & (rb >= 0) & (rb + RQ_WIN <= T.rows) & (T.generic == 0);
    // A quarter warp is one 128-bit shared-memory phase: its eight windows start in arbitrary 16-byte bank groups, two to three
    // per group.  Each lane therefore walks its window's columns ROTATED -- register slot i holds window column (i + sh) & 7
    // with sh chosen so that slot i of lane l lies in bank group (l + i) & 7: eight distinct groups in every phase.  The x
    // weights are rotated the same way once, so the arithmetic does not see it.
    const int sh = (l - (rb * pitch + cb)) & 7;
    // each pixel's six weights at its offset (0 .. 2) in the shared window; the x rows then rotated left by sh
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int ox = cs[p] - cmin, oy = rs[p] - rmin;
        ok = ok & (ox <= 2) & (oy <= 2);
        const int rot = (sh - ox) & 7; // slot i holds weight index (i + sh - ox) & 7
#pragma unroll
        for (int bit = 0; bit < 3; ++bit) {
            const bool on = (rot >> bit) & 1;
            float tmp[RQ_WIN];
#pragma unroll
            for (int k = 0; k < RQ_WIN; ++k) tmp[k] = exw[p][(k + (1 << bit)) & 7];
#pragma unroll
            for (int k = 0; k < RQ_WIN; ++k) exw[p][k] = on ? tmp[k] : exw[p][k];
        }
#pragma unroll
        for (int k = RQ_WIN - 1; k >= 0; --k) {
            const float b1 = k >= 1 ? eyw[p][k - 1] : 0.0f, b2 = k >= 2 ? eyw[p][k - 2] : 0.0f;
            eyw[p][k] = oy == 0 ? eyw[p][k] : (oy == 1 ? b1 : b2);
        }
    }
    mbar_wait(&bar, 0);
    if (TYPE != TC_F32) {
        // the tile arrived in the storage type (a quarter / half of the bytes through L2): one conversion per staged pixel
        const int n = T.rows * pitch;
#pragma unroll 4
        for (int i = tid; i < n; i += RQ_THREADS) rot_smem[i] = raw_px(raw[i], TYPE);
        __syncthreads();
        if (!live) return;
    }
    if (!ok) {
#pragma unroll 1
        for (int p = 0; p < 4; ++p)
            if (x + (p & 1) < dst.w && y + (p >> 1) < dst.h) rot_px_generic(rot_smem, pitch, T.rows, s0, t0, se[p], te[p], dst, x + (p & 1), y + (p >> 1));
        return;
    }
    float4 sum[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) sum[p] = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4* col[RQ_WIN];
#pragma unroll
    for (int i = 0; i < RQ_WIN; ++i) col[i] = rot_smem + rb * pitch + cb + ((i + sh) & 7);
#pragma unroll
    for (int j = 0; j < RQ_WIN; ++j) {
        float4 v[RQ_WIN];
#pragma unroll
        for (int i = 0; i < RQ_WIN; ++i) {
            v[i] = *col[i];
            col[i] += pitch;
        }
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            float4 r = mul4p(exw[p][0], v[0]);
#pragma unroll
            for (int i = 1; i < RQ_WIN; ++i) r = fma4p(exw[p][i], v[i], r);
            sum[p] = fma4p(eyw[p][j], r, sum[p]);
        }
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        if (x + (p & 1) < dst.w && y + (p >> 1) < dst.h) {
            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
            if (total[p] > 0.0f) o = mul4p(rcp_approx(total[p]), sum[p]);
            store_px(dst.p, dst.type, (size_t)(y + (p >> 1)) * dst.w + x + (p & 1), o);
        }
    }
}

static int filter_id(const char* name)
{
    if (!name || !*name) return -1;
    if (!strcmp(name, "lanczos3")) return 0;
    if (!strcmp(name, "blackman-harris")) return 1;
    return -2;
}

// Per-axis tap tables are a function of the geometry and the filter only: built once per (device, axis
// parameters) and kept, so a resize is one launch.  The first use synchronises its stream, so later uses
// from any stream see a finished table.
struct AxisTable { int* first = nullptr; float* wts = nullptr; float* total = nullptr; };

static int axis_table(const AxisParams& A, cudaStream_t st, AxisTable& out)
{
    static std::mutex mutex;
    static std::map<std::tuple<int, int, int, int, uint32_t, uint32_t, int>, AxisTable> cache;
    int device = 0;
    TC_CUDA(cudaGetDevice(&device));
    uint32_t sb, rb;
    memcpy(&sb, &A.scale, 4);
    memcpy(&rb, &A.ratio, 4);
    const auto key = std::make_tuple(device, A.dn, A.sn, A.id, sb, rb, A.taps);
    std::lock_guard<std::mutex> lock(mutex);
    const auto it = cache.find(key);
    if (it != cache.end()) {
        out = it->second;
        return TC_OK;
    }
    if (cache.size() >= 512) return TC_ERR_UNSUPPORTED; // caller builds the tables in its scratch instead
    AxisTable t;
    const size_t bytes = ((size_t)A.dn * (3 + A.taps)) * 4;
    void* p = nullptr;
    TC_CUDA(cudaMalloc(&p, bytes));
    t.first = (int*)p;
    t.total = (float*)p + A.dn;
    t.wts = t.total + 2 * (size_t)A.dn;
    k_resize_axis<<<(A.dn + 127) / 128, 128, 0, st>>>(A, t.first, t.wts, t.total);
    count_launch();
    TC_CUDA(cudaGetLastError());
    TC_CUDA(cudaStreamSynchronize(st));
    cache[key] = t;
    out = t;
    return TC_OK;
}

} // namespace tc

using namespace tc;

extern "C" {

int tc_resize(const tc_image* dst, const tc_image* src, const char* filter_name, float filter_width, tc_stream s)
{
    int v;
    if ((v = validate(dst, "tc_resize dst")) != TC_OK) return v;
    if ((v = validate(src, "tc_resize src")) != TC_OK) return v;
    int id = filter_id(filter_name);
    if (id == -2) return fail(TC_ERR_UNSUPPORTED, "tc_resize: filter '%s' (lanczos3 and blackman-harris are in scope)", filter_name);
    // default filter choice of ImageBufAlgo::resize
    const float wr = (float)dst->w / (float)src->w, hr = (float)dst->h / (float)src->h;
    if (id < 0) id = (wr > 1.0f || hr > 1.0f) ? 1 : 0;
    const float base = id == 0 ? 6.0f : 3.0f;
    const float fwx = filter_width > 0.0f ? filter_width : base * (wr > 1.0f ? wr : 1.0f);
    const float fwy = filter_width > 0.0f ? filter_width : base * (hr > 1.0f ? hr : 1.0f);
    AxisParams ax { dst->w, src->w, id, id == 0 ? 6.0f / fwx : 2.0f / fwx, wr, 0, 0 };
    AxisParams ay { dst->h, src->h, id, id == 0 ? 6.0f / fwy : 2.0f / fwy, hr, 0, 0 };
    ax.radi = (int)ceilf((fwx / 2.0f) / wr);
    ay.radi = (int)ceilf((fwy / 2.0f) / hr);
    ax.taps = 2 * ax.radi + 1;
    ay.taps = 2 * ay.radi + 1;
    if (ax.taps > 4096 || ay.taps > 4096) return fail(TC_ERR_UNSUPPORTED, "tc_resize: %d x %d taps", ax.taps, ay.taps);

    cudaStream_t st = (cudaStream_t)s;
    // the march: both tap counts within a band, block spans within the expanded-weight tables, tile within shared memory
    ResizeMarch M;
    M.xtaps = ax.taps;
    M.ytaps = ay.taps;
    M.sw = (int)ceilf((float)(RM_W - 1) * (float)src->w / (float)dst->w) + ax.taps + 1;
    const int span_x = (int)ceilf((float)(RM_B - 1) * (float)src->w / (float)dst->w) + ax.taps + 1;
    const int span_y = (int)ceilf((float)(RM_B - 1) * (float)src->h / (float)dst->h) + ay.taps + 1;
    M.ring = ay.taps <= 16 ? 48 : 64;
    M.wspan = std::min(RM_SPAN, (std::max(span_x, span_y) + 3) & ~3);
    const size_t march_smem = resize_march_smem(M.sw, M.ring, M.wspan);
    static const int march_occ = [] {
        const char* e = getenv("TOUCAN_RESIZE_OCC"); // 2: always the register-prefetch variant (two CTAs per SM)
        return e && *e ? atoi(e) : 3;
    }();
    // three CTAs per SM where they fit; a one-span tile (enlargements) holds 4 pixels per thread in flight, measured faster with the prefetch
    const bool three = march_occ >= 3 && march_smem <= 74 * 1024 && M.sw > 32;
    const bool march = ax.taps <= RM_BAND && ay.taps <= RM_BAND && span_x <= RM_SPAN && span_y <= RM_SPAN && M.sw <= 128 && march_smem <= 99 * 1024;
    // per-axis tables: cached per geometry; built in the call's scratch only if the cache is full
    AxisTable tx, ty;
    const bool cached = axis_table(ax, st, tx) == TC_OK && axis_table(ay, st, ty) == TC_OK;
    // scratch: [tmp (dw x sh float4, two-pass path only) | xf | yf | xtot | ytot | xw | yw (uncached tables only)]
    const bool direct = !march && (size_t)ax.taps * ay.taps >= 4096 && (double)dst->w * dst->h * ax.taps * ay.taps <= 4.0e9;
    const size_t n_tmp = (march || direct) ? 0 : (size_t)dst->w * src->h * 4;
    const size_t n_int = cached ? 0 : (size_t)dst->w + dst->h;
    const size_t n_tot = cached ? 0 : 2 * ((size_t)dst->w + dst->h);
    const size_t n_w = cached ? 0 : (size_t)dst->w * ax.taps + (size_t)dst->h * ay.taps;
    void* scratch = nullptr;
    if (n_tmp + n_int + n_tot + n_w) TC_CUDA(cudaMallocAsync(&scratch, (n_tmp + n_int + n_tot + n_w) * 4, st));
    float4* tmp = (float4*)scratch;
    int* xf = (int*)((float*)scratch + n_tmp);
    int* yf = xf + dst->w;
    float* xt = (float*)(yf + dst->h);
    float* yt = xt + 2 * (size_t)dst->w;
    float* xw = yt + 2 * (size_t)dst->h;
    float* yw = xw + (size_t)dst->w * ax.taps;
    if (cached) {
        xf = tx.first, xw = tx.wts, xt = tx.total;
        yf = ty.first, yw = ty.wts, yt = ty.total;
    } else {
        k_resize_axis<<<(dst->w + 127) / 128, 128, 0, st>>>(ax, xf, xw, xt);
        k_resize_axis<<<(dst->h + 127) / 128, 128, 0, st>>>(ay, yf, yw, yt);
        count_launch(2);
    }
    // float32 frames on a 16-byte aligned base, up to 16 taps per axis: warp-private rings behind TMA-staged bands
    static const bool wp_on = [] {
        const char* e = getenv("TOUCAN_RESIZE_WP");
        return !(e && *e == '0');
    }();
    bool wp_done = false;
    // (reductions only: enlarging, a band completes many rows and the march's CTA-wide column pass is the faster one -- 0.072 against 0.079 ms at 2K -> 4K)
    if (march && wp_on && src->type == TC_F32 && dst->type == TC_F32 && dst->h <= src->h && tma_available() && ((uintptr_t)src->data & 15) == 0 && ay.taps <= 16 &&
        ax.taps <= WP_SPAN) {
        const float rows_per_band = (float)WP_BAND * (float)dst->h / (float)src->h;
        const int brl = rows_per_band <= 8.5f ? 2 : (rows_per_band <= 16.5f ? 4 : 8);
        ResizeWp Q;
        Q.xtaps = ax.taps;
        Q.ytaps = ay.taps;
        Q.sw = (int)ceilf(31.0f * (float)src->w / (float)dst->w) + ax.taps + 1;
        const int wspan_x = (int)ceilf((float)(RM_B - 1) * (float)src->w / (float)dst->w) + ax.taps + 1;
        const int wspan_y = (int)ceilf((float)(brl - 1) * (float)src->h / (float)dst->h) + ay.taps + 1;
        // one sub-strip (32 columns, 4 warps, one staged tile) per CTA and four CTAs per SM: 0.073 ms at 4K -> 2K against 0.077 ms
        // with two sub-strips per CTA and two CTAs per SM (more copies in flight per SM for the same warps)
        constexpr int wp_nsub = 1;
        auto smem_for = [&](int stages) {
            return brl == 2 ? resize_wp_smem<2, 1>(Q.sw, stages) : (brl == 4 ? resize_wp_smem<4, 1>(Q.sw, stages) : resize_wp_smem<8, 1>(Q.sw, stages));
        };
        // two CTAs per SM; two staging buffers where they fit next to the rings, else one (the other CTA covers the copy).  Measured
        // at 4K -> 2K: one buffer at two CTAs per SM 0.075 ms, four buffers at ONE CTA per SM 0.126 ms -- warps, not copies in flight
        Q.stages = smem_for(2) <= (wp_nsub == 2 ? 112 : 56) * 1024 ? 2 : 1;
        const size_t wp_smem = smem_for(Q.stages);
        // as many CTAs per SM as the shared memory holds (five at 4K -> 2K: 0.073 -> 0.066 ms against a grid sized for four -- the
        // kernel is latency bound, every further warp counts)
        const int per_sm = std::max(1, std::min(wp_nsub == 2 ? 2 : 8, (int)((size_t)227 * 1024 / (wp_smem + 1024))));
        CUtensorMap map;
        if (wspan_x <= WP_SPAN && wspan_y <= WP_SPAN && 2 * (Q.sw | 1) <= 256 && wp_smem <= 112 * 1024 &&
            tensor_map_2d(src->data, 8, 2ull * src->w, (unsigned long long)src->h, 16ull * src->w, 2u * (Q.sw | 1), (unsigned)WP_BAND, &map) == TC_OK) {
            const int strip_w = 32 * wp_nsub;
            const int strips = (dst->w + strip_w - 1) / strip_w;
            int chunks = std::max(1, per_sm * num_sms() / strips);
            chunks = std::min(chunks, std::max(1, src->h / (8 * ay.taps)));
            chunks = std::max(1, std::min(chunks, dst->h));
            Q.ch = (dst->h + chunks - 1) / chunks;
            const int max_rows = std::max(1, (int)((double)(WP_BAND * (RM_MAXBANDS - 2) - ay.taps) * dst->h / src->h));
            Q.ch = std::min(Q.ch, max_rows);
            const dim3 grid(strips, (dst->h + Q.ch - 1) / Q.ch);
            int e = TC_OK;
#define TC_WP2(B, N)                                                                                                            \
    e = ensure_dynamic_smem((const void*)k_resize_wp<B, N>, wp_smem);                                                            \
    if (e == TC_OK) k_resize_wp<B, N><<<grid, 128 * N, wp_smem, st>>>(view(src), view(dst), xf, xw, xt, yf, yw, yt, Q, map);
#define TC_WP(B) TC_WP2(B, 1)
            if (brl == 2) { TC_WP(2) } else if (brl == 4) { TC_WP(4) } else { TC_WP(8) }
#undef TC_WP2
#undef TC_WP
            if (e != TC_OK) {
                if (scratch) cudaFreeAsync(scratch, st);
                return e;
            }
            count_launch(1);
            wp_done = true;
        }
    }
    if (wp_done) {
    } else if (march) {
        // chunks of destination rows: one wave of two / three CTAs per SM, warm-up rows (ytaps per chunk) kept below ~1/8 of a chunk
        const int strips = (dst->w + RM_W - 1) / RM_W;
        int chunks = std::max(1, (three ? 3 : 2) * num_sms() / strips);
        chunks = std::min(chunks, std::max(1, src->h / (8 * ay.taps)));
        chunks = std::max(1, std::min(chunks, dst->h));
        M.ch = (dst->h + chunks - 1) / chunks;
        // a chunk walks at most RM_MAXBANDS bands of source rows
        const int max_rows = std::max(1, (int)((double)(RM_BAND * (RM_MAXBANDS - 2) - ay.taps) * dst->h / src->h));
        M.ch = std::min(M.ch, max_rows);
        const dim3 grid(strips, (dst->h + M.ch - 1) / M.ch);
        int e = TC_OK;
#define TC_RM3(TY, CS, PF)                                                                                                          \
    e = ensure_dynamic_smem((const void*)k_resize_march<TY, CS, PF>, march_smem);                                                    \
    if (e == TC_OK) k_resize_march<TY, CS, PF><<<grid, RM_THREADS, march_smem, st>>>(view(src), view(dst), xf, xw, xt, yf, yw, yt, M);
#define TC_RM2(TY, CS)                                                                                                              \
    if (three) { TC_RM3(TY, CS, false) } else { TC_RM3(TY, CS, true) }
#define TC_RM(TY)                                                                                                                   \
    case TY:                                                                                                                        \
        if (M.sw <= 32) { TC_RM2(TY, 1) }                                                                                           \
        else if (M.sw <= 64) { TC_RM2(TY, 2) }                                                                                      \
        else if (M.sw <= 96) { TC_RM2(TY, 3) }                                                                                      \
        else { TC_RM2(TY, 4) }                                                                                                      \
        break;
        switch (src->type) {
            TC_RM(TC_U8)
            TC_RM(TC_U16)
            TC_RM(TC_F16)
            TC_RM(TC_F32)
        }
#undef TC_RM
#undef TC_RM3
#undef TC_RM2
#undef TC_RM
        if (e != TC_OK) {
            if (scratch) cudaFreeAsync(scratch, st);
            return e;
        }
        count_launch(1);
    } else if (direct) {
        const int n = dst->w * dst->h;
#define TC_RD(TY)                                                                                                                  \
    case TY: k_resize_direct<TY><<<(n + 127) / 128, 128, 0, st>>>(view(src), view(dst), xf, xw, xt, yf, yw, yt, ax.taps, ay.taps); break;
        switch (src->type) {
            TC_RD(TC_U8)
            TC_RD(TC_U16)
            TC_RD(TC_F16)
            TC_RD(TC_F32)
        }
#undef TC_RD
        count_launch(1);
    } else {
        dim3 block(32, 8);
        // row pass: staged tile when a useful number of destination columns fits next to their footprint, else the gather
        RowsTile R;
        R.xtaps = ax.taps;
        R.cols = 127;
        R.xw = (int)floorf((float)(R.cols - ax.taps - 1) * (float)dst->w / (float)src->w) + 1;
        if (R.xw > 64) R.xw = 64;
        const size_t rows_smem = (size_t)RR_ROWS * (R.cols | 1) * sizeof(float4) + ((size_t)R.xw * ax.taps + R.xw) * sizeof(float);
        if (R.xw >= 4 && ax.taps + 2 <= R.cols && rows_smem <= 72 * 1024) {
            const dim3 rgrid((dst->w + R.xw - 1) / R.xw, (src->h + RR_ROWS - 1) / RR_ROWS);
            int e = TC_OK;
#define TC_RR(TY)                                                                                                              \
    case TY:                                                                                                                   \
        e = ensure_dynamic_smem((const void*)k_resize_rows_tile<TY>, rows_smem);                                               \
        if (e == TC_OK) k_resize_rows_tile<TY><<<rgrid, RR_THREADS, rows_smem, st>>>(view(src), tmp, dst->w, xf, xw, xt, R);   \
        break;
            switch (src->type) {
                TC_RR(TC_U8)
                TC_RR(TC_U16)
                TC_RR(TC_F16)
                TC_RR(TC_F32)
            }
#undef TC_RR
            if (e != TC_OK) {
                if (scratch) cudaFreeAsync(scratch, st);
                return e;
            }
        } else
        k_resize_rows<<<dim3((dst->w + 31) / 32, (src->h + 7) / 8), block, 0, st>>>(view(src), tmp, dst->w, xf, xw, xt, ax.taps);
        k_resize_cols<<<dim3((dst->w + 31) / 32, (dst->h + 7) / 8), block, 0, st>>>(tmp, src->h, view(dst), yf, yw, yt, ay.taps);
        count_launch(2);
    }
    int st_launch = check_cuda(cudaGetLastError(), "k_resize launch");
    if (scratch) cudaFreeAsync(scratch, st);
    return st_launch;
}

int tc_rotate(const tc_image* dst, const tc_image* src, float angle, const char* filter_name, float filter_width, tc_stream s)
{
    int v;
    if ((v = validate(dst, "tc_rotate dst")) != TC_OK) return v;
    if ((v = validate(src, "tc_rotate src")) != TC_OK) return v;
    int id = filter_id(filter_name);
    if (id == -2) return fail(TC_ERR_UNSUPPORTED, "tc_rotate: filter '%s' (lanczos3 and blackman-harris are in scope)", filter_name);
    if (id < 0) id = 0; // warp's default filter is lanczos3
    RotParams P;
    P.id = id;
    P.fw = filter_width > 0.0f ? filter_width : (id == 0 ? 6.0f : 3.0f);
    P.fscale = id == 0 ? 6.0f / P.fw : 2.0f / P.fw;
    if (P.fw + 2.0f > (float)ROT_MAX_TAPS) return fail(TC_ERR_UNSUPPORTED, "tc_rotate: filter width %g too large", P.fw);
    // M = T(-c) * R(angle) * T(c) (Imath row-vector convention), centre of the source window
    const float cx = 0.5f * (float)(0 + src->w), cy = 0.5f * (float)(0 + src->h);
    const float c = cosf(angle), sn = sinf(angle);
    volatile float m20 = -cx * c + -cy * -sn;
    volatile float m21 = -cx * sn + -cy * c;
    m20 = m20 + cx;
    m21 = m21 + cy;
    const float m00 = c, m01 = sn, m10 = -sn, m11 = c;
    // Imath Matrix33::inverse(), affine case
    volatile float r = m00 * m11 - m10 * m01;
    volatile float i00 = m11 / r, i01 = -m01 / r, i10 = -m10 / r, i11 = m00 / r;
    volatile float i20 = -m20 * i00 - m21 * i10;
    volatile float i21 = -m20 * i01 - m21 * i11;
    const float mi[9] = { i00, i01, 0.0f, i10, i11, 0.0f, i20, i21, 1.0f };
    memcpy(P.m, mi, sizeof(mi));
    dim3 block(32, 8), grid((dst->w + 31) / 32, (dst->h + 7) / 8);
    const bool narrow = P.fw + 2.0f <= 8.0f;
    const Img vs = view(src), vd = view(dst);
    cudaStream_t st = (cudaStream_t)s;
    // four pixels per thread: plain rotations with the default lanczos3, tile moved by the copy engine in the storage type
    {
        static const int rot_quad_mode = [] {
            const char* e = getenv("TOUCAN_ROTATE_QUAD");
            return (e && *e == '0') ? 0 : ((e && !strcmp(e, "generic")) ? 2 : 1);
        }();
        const bool rot_quad = rot_quad_mode != 0;
        const float ds = fmaxf(1.0f, fmaxf(fabsf(P.m[0]), fabsf(P.m[3]))), dt = fmaxf(1.0f, fmaxf(fabsf(P.m[1]), fabsf(P.m[4])));
        const float spx = fabsf(P.m[0]) + fabsf(P.m[3]), spy = fabsf(P.m[1]) + fabsf(P.m[4]);
        if (rot_quad && id == 0 && P.fw == 6.0f && ds == 1.0f && dt == 1.0f && spx <= 1.9f && spy <= 1.9f && src->type == dst->type) {
            RotTile T;
            T.generic = rot_quad_mode == 2 ? 1 : 0;
            const float ex = (float)(RQ_BW - 1), ey = (float)(RQ_BH - 1);
            T.cols = (int)ceilf(ex * fabsf(P.m[0]) + ey * fabsf(P.m[3]) + 6.0f) + 7 + (src->type == TC_U8 ? 3 : (src->type == TC_F32 ? 0 : 1));
            T.rows = (int)ceilf(ex * fabsf(P.m[1]) + ey * fabsf(P.m[4]) + 6.0f) + 7;
            T.smn_off = fminf(0.0f, ex * P.m[0]) + fminf(0.0f, ey * P.m[3]);
            T.smx_off = fmaxf(0.0f, ex * P.m[0]) + fmaxf(0.0f, ey * P.m[3]);
            T.tmn_off = fminf(0.0f, ex * P.m[1]) + fminf(0.0f, ey * P.m[4]);
            T.tmx_off = fmaxf(0.0f, ex * P.m[1]) + fmaxf(0.0f, ey * P.m[4]);
            // row pitch: the kernel's per-lane column rotation removes the bank conflicts for any pitch; a box row of the copy
            // is a multiple of 16 bytes in the storage type
            const int pal = src->type == TC_U8 ? 4 : (src->type == TC_F32 ? 1 : 2);
            T.pitch = (T.cols + pal - 1) / pal * pal;
            const size_t tile = ((size_t)T.rows * T.pitch * sizeof(float4) + 127) / 128 * 128;
            const size_t px_bytes = src->type == TC_U8 ? 4 : (src->type == TC_F32 ? 16 : 8);
            size_t smem = tile;
            T.raw_off = 0;
            CUtensorMap map;
            memset(&map, 0, sizeof(map));
            if (smem <= 72 * 1024) {
                static const bool rot_tma = [] {
                    const char* e = getenv("TOUCAN_ROTATE_TMA");
                    return !(e && *e == '0');
                }();
                // the copy engine wants the source base and row pitch on 16-byte boundaries
                bool tma = rot_tma && tma_available() && ((uintptr_t)src->data & 15) == 0 && ((size_t)src->w * px_bytes) % 16 == 0 && T.pitch <= 128 && T.rows <= 256;
                if (tma && src->type == TC_F32)
                    tma = tensor_map_2d(src->data, 8, 2ull * src->w, (unsigned long long)src->h, 16ull * src->w, 2u * T.pitch, (unsigned)T.rows, &map) == TC_OK;
                else if (tma) {
                    tma = tensor_map_2d(src->data, (int)px_bytes, (unsigned long long)src->w, (unsigned long long)src->h, px_bytes * src->w, (unsigned)T.pitch,
                                        (unsigned)T.rows, &map) == TC_OK;
                    if (tma) {
                        T.raw_off = (int)tile;
                        smem = tile + (size_t)T.rows * T.pitch * px_bytes;
                    }
                }
                const dim3 qgrid((dst->w + RQ_BW - 1) / RQ_BW, (dst->h + RQ_BH - 1) / RQ_BH);
                int e = TC_OK;
#define TC_ROTQ(TY)                                                                       \
    case TY:                                                                              \
        e = ensure_dynamic_smem((const void*)k_rotate_quad<TY>, smem);                    \
        if (e == TC_OK) k_rotate_quad<TY><<<qgrid, RQ_THREADS, smem, st>>>(vd, P, T, map); \
        break;
                if (tma) {
                    switch (dst->type) {
                        TC_ROTQ(TC_U8)
                        TC_ROTQ(TC_U16)
                        TC_ROTQ(TC_F16)
                        TC_ROTQ(TC_F32)
                    }
                    if (e != TC_OK) return e;
                    count_launch();
                    return check_cuda(cudaGetLastError(), "k_rotate_quad launch");
                }
#undef TC_ROTQ
                // (no copy engine for this frame -- unaligned base or row pitch: the one-pixel-per-thread kernels below)
            }
        }
    }
    // shared-memory tile path: default-width filters whose block footprint fits 64 KB
    {
        const float ds = fmaxf(1.0f, fmaxf(fabsf(P.m[0]), fabsf(P.m[3]))), dt = fmaxf(1.0f, fmaxf(fabsf(P.m[1]), fabsf(P.m[4])));
        RotTile T;
        T.generic = 0;
        T.cols = (int)ceilf((RTT - 1) * (fabsf(P.m[0]) + fabsf(P.m[3])) + ds * P.fw) + 5;
        T.rows = (int)ceilf((RTT - 1) * (fabsf(P.m[1]) + fabsf(P.m[4])) + dt * P.fw) + 5;
        const float edge = (float)(RTT - 1);
        T.smn_off = fminf(0.0f, edge * P.m[0]) + fminf(0.0f, edge * P.m[3]);
        T.smx_off = fmaxf(0.0f, edge * P.m[0]) + fmaxf(0.0f, edge * P.m[3]);
        T.tmn_off = fminf(0.0f, edge * P.m[1]) + fminf(0.0f, edge * P.m[4]);
        T.tmx_off = fmaxf(0.0f, edge * P.m[1]) + fmaxf(0.0f, edge * P.m[4]);
        // Row pitch: a quarter warp (one 128-bit shared-memory phase) is a 4 x 2 block of destination pixels whose footprints
        // start at (a m0 + b m3, a m1 + b m4) from the first; pick the pitch whose 16-byte slots collide least over a few
        // sub-pixel phases of that first position.
        int best = 1 << 30;
        T.pitch = T.cols;
        for (int pitch = T.cols; pitch < T.cols + 8; ++pitch) {
            int cost = 0;
            for (int ph = 0; ph < 16; ++ph) {
                const float sf = 0.25f * (float)(ph & 3), tf = 0.25f * (float)(ph >> 2);
                // wavefronts of one phase = the most DISTINCT 16-byte slots that share a bank group
                int slots[8], worst = 0;
                for (int k = 0; k < 8; ++k) {
                    const float a = (float)(k & 3), b = (float)(k >> 2);
                    slots[k] = (int)floorf(tf + 64.0f + a * P.m[1] + b * P.m[4]) * pitch + (int)floorf(sf + 64.0f + a * P.m[0] + b * P.m[3]);
                }
                for (int g = 0; g < 8; ++g) {
                    int n = 0;
                    for (int k = 0; k < 8; ++k) {
                        if ((slots[k] & 7) != g) continue;
                        bool seen = false;
                        for (int q = 0; q < k; ++q) seen = seen || slots[q] == slots[k];
                        n += seen ? 0 : 1;
                    }
                    worst = n > worst ? n : worst;
                }
                cost += worst;
            }
            if (cost < best) {
                best = cost;
                T.pitch = pitch;
            }
        }
        const size_t smem = (size_t)T.rows * T.pitch * sizeof(float4);
        if (narrow && smem <= 64 * 1024) {
            const dim3 tgrid((dst->w + RTT - 1) / RTT, (dst->h + RTT - 1) / RTT);
            int e = TC_OK;
            // float32 sources on a 16-byte aligned base: the box through the copy engine
            CUtensorMap map;
            memset(&map, 0, sizeof(map));
            static const bool rot_tma = [] {
                const char* e = getenv("TOUCAN_ROTATE_TMA");
                return !(e && *e == '0');
            }();
            const bool tma = rot_tma && src->type == TC_F32 && tma_available() && ((uintptr_t)src->data & 15) == 0 && 2 * T.pitch <= 256 && T.rows <= 256 &&
                             tensor_map_2d(src->data, 8, 2ull * src->w, (unsigned long long)src->h, 16ull * src->w, 2u * T.pitch, (unsigned)T.rows, &map) == TC_OK;
#define TC_ROTT2(TY, F, TM)                                                                                             \
    e = ensure_dynamic_smem((const void*)k_rotate_tile<8, F, TY, TM>, smem);                                            \
    if (e == TC_OK) k_rotate_tile<8, F, TY, TM><<<tgrid, RTT * RTT, smem, st>>>(vs, vd, P, T, map);
#define TC_ROTT(TY)                                                                                                     \
    case TY:                                                                                                            \
        if (id == 0) { TC_ROTT2(TY, 0, false) } else { TC_ROTT2(TY, 1, false) }                                         \
        break;
            if (tma) {
                if (id == 0) { TC_ROTT2(TC_F32, 0, true) } else { TC_ROTT2(TC_F32, 1, true) }
            } else
            switch (src->type) {
                TC_ROTT(TC_U8)
                TC_ROTT(TC_U16)
                TC_ROTT(TC_F16)
                TC_ROTT(TC_F32)
            }
#undef TC_ROTT2
#undef TC_ROTT
            if (e != TC_OK) return e;
            count_launch();
            return check_cuda(cudaGetLastError(), "k_rotate_tile launch");
        }
    }
#define TC_ROT(T)                                                                         \
    case T:                                                                               \
        if (id == 0) {                                                                    \
            if (narrow) k_rotate<8, 0, T><<<grid, block, 0, st>>>(vs, vd, P);             \
            else k_rotate<16, 0, T><<<grid, block, 0, st>>>(vs, vd, P);                   \
        } else {                                                                          \
            if (narrow) k_rotate<8, 1, T><<<grid, block, 0, st>>>(vs, vd, P);             \
            else k_rotate<16, 1, T><<<grid, block, 0, st>>>(vs, vd, P);                   \
        }                                                                                 \
        break;
    switch (src->type) {
        TC_ROT(TC_U8)
        TC_ROT(TC_U16)
        TC_ROT(TC_F16)
        TC_ROT(TC_F32)
    }
#undef TC_ROT
    count_launch();
    return check_cuda(cudaGetLastError(), "k_rotate launch");
}

} // extern "C"
