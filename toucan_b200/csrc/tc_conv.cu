// tc_conv.cu -- toucan:Blur and toucan:UnsharpMask: the reference's 2-D Gaussian
// convolve (O(k^2) taps per pixel, plugins/FilterPlugin.cpp:188-202 / :607-618) as a
// separable shared-memory tiled stencil, one read of the source and one write of the
// result per pixel.  Edge handling = clamp (OIIO convolve uses WrapClamp), so border
// tiles are staged with clamped coordinates rather than zero-filling bulk copies.
//
// Tolerance op (<= 1e-5 abs / +-1 LSB), so FMA contraction is allowed in this file.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "tc_tile.cuh"
#include "tc_tma.cuh"

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


// =====================================================================================================
// k_gauss_tma: Blur / UnsharpMask with the tile moved by the TMA engine IN ITS STORAGE TYPE (a u8 tile is a
// quarter of the float one: a quarter of the shared-memory bytes on the way in and out of the row pass, and a
// quarter of the LDS instructions -- one 16-byte read is four pixels), persistent CTAs that filter tile i while
// tiles i+1 / i+2 are in flight, warp-private strips between the two passes (no CTA-wide barrier; see tc_stack.cu
// for the scheme).  The box is the smallest width >= 56 + (pixels one row item reads) whose row pitch is an odd
// multiple of 16 bytes, so the row-per-lane 16-byte reads are bank-conflict free for every type.
// 8/16-bit sources: the 1/255 (1/65535) of OIIO's load conversion is folded into the row-pass weights (the taps see
// integers); the unsharp epilogue converts its centre pixel exactly.
template <int R, int TYPE, int NBUF = 2>
struct GtTile {
    static constexpr int TH = 24, PX = 6;
    static constexpr int B = TYPE == TC_U8 ? 4 : (TYPE == TC_F32 ? 16 : 8); // bytes per pixel
    static constexpr int VPX = 16 / B;                                      // pixels per 16-byte vector
    // TMA wants the first byte of a box row 16-byte aligned: the box starts RA >= R pixels left of the tile, RA a multiple of
    // the pixels per 16 bytes (found the hard way: an unaligned start is an "illegal instruction" at the copy)
    static constexpr int RA = (R + VPX - 1) / VPX * VPX;
    static constexpr int D = RA - R;                                        // tap t of output k reads window pixel k + t + D
    static constexpr int NPX = (ST_HB + 2 * R + D + VPX - 1) / VPX * VPX;   // pixels one row item reads
    static constexpr int NV = NPX / VPX;
    static constexpr int IH = TH + 2 * R;
    static constexpr int MINW = ST_TW - ST_HB + NPX;
    static constexpr int box_width()
    {
        int w = MINW;
        while ((w * B) % 32 != 16) ++w;
        return w;
    }
    static constexpr int BW = box_width();
    static constexpr int PITCH = BW * B;
    static constexpr int IN_BYTES = (IH * PITCH + 127) / 128 * 128;
    static constexpr int MP = ST_HB + 1;
    static constexpr int MID_BYTES = 8 * IH * MP * 16;
    static constexpr size_t SMEM = NBUF * IN_BYTES + MID_BYTES + 64;
    static constexpr int ELEM = B == 4 ? 4 : 8; // tensor-map element size
    static constexpr unsigned TILE_BYTES = (unsigned)(IH * PITCH);
};

template <int TYPE>
__device__ __forceinline__ void unpack_vec(uint4 v, float4* px)
{
    if (TYPE == TC_F32) {
        px[0] = make_float4(__uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z), __uint_as_float(v.w));
    } else if (TYPE == TC_F16) {
        const float2 a = __half22float2(*reinterpret_cast<__half2*>(&v.x)), b = __half22float2(*reinterpret_cast<__half2*>(&v.y));
        const float2 c = __half22float2(*reinterpret_cast<__half2*>(&v.z)), d = __half22float2(*reinterpret_cast<__half2*>(&v.w));
        px[0] = make_float4(a.x, a.y, b.x, b.y);
        px[1] = make_float4(c.x, c.y, d.x, d.y);
    } else if (TYPE == TC_U16) {
        px[0] = make_float4((float)(v.x & 0xffffu), (float)(v.x >> 16), (float)(v.y & 0xffffu), (float)(v.y >> 16));
        px[1] = make_float4((float)(v.z & 0xffffu), (float)(v.z >> 16), (float)(v.w & 0xffffu), (float)(v.w >> 16));
    } else {
        const unsigned w[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
        for (int i = 0; i < 4; ++i)
            px[i] = make_float4((float)(w[i] & 0xffu), (float)((w[i] >> 8) & 0xffu), (float)((w[i] >> 16) & 0xffu), (float)(w[i] >> 24));
    }
}
// one pixel of the staged tile, converted exactly like a load from global memory
template <int TYPE>
__device__ __forceinline__ float4 staged_px(const unsigned char* tile, int byte_offset)
{
    if (TYPE == TC_F32) return *reinterpret_cast<const float4*>(tile + byte_offset);
    if (TYPE == TC_U8) {
        const uchar4 v = *reinterpret_cast<const uchar4*>(tile + byte_offset);
        return make_float4(u8_to_f(v.x), u8_to_f(v.y), u8_to_f(v.z), u8_to_f(v.w));
    }
    const uint2 raw = *reinterpret_cast<const uint2*>(tile + byte_offset);
    if (TYPE == TC_U16) return make_float4(u16_to_f(raw.x & 0xffffu), u16_to_f(raw.x >> 16), u16_to_f(raw.y & 0xffffu), u16_to_f(raw.y >> 16));
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&raw.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
    return make_float4(a.x, a.y, b.x, b.y);
}

// The epilogue: simple same-position graph nodes that follow the convolution node, applied to the finished pixel in registers.
// Explicit round-to-nearest operations (no contraction): bit-identical to the pointwise kernels' functors (tc_pointwise.cu).
template <int TYPE>
__device__ __forceinline__ float4 epi_apply(const EpiSteps& e, float4 p)
{
    p = quant4(p, TYPE); // the convolution node's own typed output
    for (int i = 0; i < e.n; ++i) {
        switch (e.kind[i]) {
        case TC_CHAIN_INVERT:
            p.x = __fadd_rn(__fmul_rn(p.x, -1.0f), 1.0f);
            p.y = __fadd_rn(__fmul_rn(p.y, -1.0f), 1.0f);
            p.z = __fadd_rn(__fmul_rn(p.z, -1.0f), 1.0f);
            break;
        case TC_CHAIN_PREMULT:
            if (p.w != 1.0f) {
                p.x = __fmul_rn(p.x, p.w);
                p.y = __fmul_rn(p.y, p.w);
                p.z = __fmul_rn(p.z, p.w);
            }
            break;
        case TC_CHAIN_POW:
            if (e.v[i] == 2.0f) p = make_float4(__fmul_rn(p.x, p.x), __fmul_rn(p.y, p.y), __fmul_rn(p.z, p.z), __fmul_rn(p.w, p.w));
            else p = make_float4(__fmul_rn(__fmul_rn(p.x, p.x), p.x), __fmul_rn(__fmul_rn(p.y, p.y), p.y), __fmul_rn(__fmul_rn(p.z, p.z), p.z),
                                 __fmul_rn(__fmul_rn(p.w, p.w), p.w));
            break;
        default: { // TC_CHAIN_OVER_FILL
            if (e.premult[i]) {
                if (p.w != 1.0f) {
                    p.x = __fmul_rn(p.x, p.w);
                    p.y = __fmul_rn(p.y, p.w);
                    p.z = __fmul_rn(p.z, p.w);
                }
                p = quant4(p, TYPE);
            }
            const float oma = __fsub_rn(1.0f, fminf(fmaxf(p.w, 0.0f), 1.0f));
            p = make_float4(__fadd_rn(p.x, __fmul_rn(oma, e.color[i][0])), __fadd_rn(p.y, __fmul_rn(oma, e.color[i][1])),
                            __fadd_rn(p.z, __fmul_rn(oma, e.color[i][2])), __fadd_rn(p.w, __fmul_rn(oma, e.color[i][3])));
            break;
        }
        }
        if (i + 1 < e.n) p = quant4(p, TYPE);
    }
    return p;
}

struct GtParams {
    int w, h, tiles_x, n_tiles;
    ConvParams c;
    EpiSteps epi;
    alignas(64) CUtensorMap map;
};

template <int R, int TYPE, bool UNSHARP, int NBUF, int OCC, bool EPI>
__global__ void __launch_bounds__(ST_THREADS, OCC) k_gauss_tma(void* __restrict__ dstp, const __grid_constant__ GtParams Q)
{
    using T = GtTile<R, TYPE, NBUF>;
    constexpr int PX = T::PX;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* buf0 = smem_raw;
    unsigned char* buf1 = smem_raw + (NBUF - 1) * T::IN_BYTES; // (== buf0 with one staging buffer)
    float4* mid = (float4*)(smem_raw + NBUF * T::IN_BYTES);
    uint64_t* full = (uint64_t*)(smem_raw + NBUF * T::IN_BYTES + T::MID_BYTES); // [2]
    int* released = (int*)(full + 2);                                           // [2]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int stride = gridDim.x;

    auto issue = [&](int s, int tile) {
        const int tx = tile % Q.tiles_x, ty = tile / Q.tiles_x;
        mbar_expect_tx(&full[s], T::TILE_BYTES);
        tma_load_tile(s ? buf1 : buf0, &Q.map, (tx * ST_TW - T::RA) * (T::B / T::ELEM), ty * T::TH - R, &full[s]);
    };
    if (threadIdx.x == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        released[0] = released[1] = 0;
        mbar_init_fence();
        if ((int)blockIdx.x < Q.n_tiles) issue(0, blockIdx.x);
        if (NBUF > 1 && (int)blockIdx.x + stride < Q.n_tiles) issue(1, blockIdx.x + stride);
    }
    __syncthreads();

    // taps centred on R; the row pass carries the load conversion's scale for integer storage
    float g[2 * R + 1], gr[2 * R + 1];
    const float in_scale = TYPE == TC_U8 ? 1.0f / 255.0f : (TYPE == TC_U16 ? 1.0f / 65535.0f : 1.0f);
#pragma unroll
    for (int t = 0; t <= 2 * R; ++t) {
        const int sft = t - R + Q.c.R;
        g[t] = (sft >= 0 && sft <= 2 * Q.c.R) ? Q.c.g[sft] : 0.0f;
        gr[t] = g[t] * in_scale;
    }
    const int xb = warp;
    float4* strip = mid + warp * (T::IH * T::MP);
    const int col = lane % ST_HB, seg = lane / ST_HB;

    int i = 0;
    for (int tile = blockIdx.x; tile < Q.n_tiles; tile += stride, ++i) {
        const int s = NBUF > 1 ? (i & 1) : 0;
        unsigned char* in = s ? buf1 : buf0;
        const int tx = tile % Q.tiles_x, ty = tile / Q.tiles_x;
        const int x0 = tx * ST_TW, y0 = ty * T::TH;
        mbar_wait(&full[s], (NBUF > 1 ? (i >> 1) : i) & 1);
        const bool border = x0 - T::RA < 0 || y0 - R < 0 || x0 - T::RA + T::BW > Q.w || y0 + T::TH + R > Q.h;
        if (border) {
            // cells outside the frame <- the clamped in-frame cell of the same tile (TMA filled them with zeros)
            for (int c = threadIdx.x; c < T::IH * T::BW; c += ST_THREADS) {
                const int r = c / T::BW, cc = c - r * T::BW;
                const int sx = x0 - T::RA + cc, sy = y0 - R + r;
                const int cx = min(max(sx, 0), Q.w - 1), cy = min(max(sy, 0), Q.h - 1);
                if (cx != sx || cy != sy) {
                    const unsigned char* from = in + (cy - (y0 - R)) * T::PITCH + (cx - (x0 - T::RA)) * T::B;
                    unsigned char* to = in + r * T::PITCH + cc * T::B;
                    if (T::B == 16) *reinterpret_cast<uint4*>(to) = *reinterpret_cast<const uint4*>(from);
                    else if (T::B == 8) *reinterpret_cast<uint2*>(to) = *reinterpret_cast<const uint2*>(from);
                    else *reinterpret_cast<unsigned*>(to) = *reinterpret_cast<const unsigned*>(from);
                }
            }
            __syncthreads();
        }
        __syncwarp();
        // ---- rows: lane = source row, 8 outputs from NV 16-byte reads ----
        // Up to 32 staged rows: one lane per row, 8 outputs each.  More rows (R >= 6): a lane per HALF row (4 outputs), so that
        // 2 IH items fill the warp's iterations (IH = 42: 3 iterations of 4 outputs instead of 2 of 8, a quarter fewer FMA slots).
        // Every lane runs every iteration on a clamped item and only the store is predicated: a partially active warp costs the
        // same issue slots.  Lanes on consecutive rows with a 16 (mod 128)-byte pitch: conflict-free either way.
        constexpr int HO = T::IH <= 32 ? ST_HB : ST_HB / 2;   // outputs per item
        constexpr int ITEMS = T::IH * (ST_HB / HO);
        constexpr int NVI = ((HO + 2 * R + T::D + T::VPX - 1) / T::VPX);  // 16-byte reads per item
        for (int i0 = 0; i0 < ITEMS; i0 += 32) {
            const int item = min(i0 + lane, ITEMS - 1);
            const int half = item / T::IH, row = item - half * T::IH;
            const uint4* p = reinterpret_cast<const uint4*>(in + row * T::PITCH + (xb * ST_HB + half * HO) * T::B);
            float4 o[HO];
#pragma unroll
            for (int v = 0; v < NVI; ++v) {
                float4 px[T::VPX];
                unpack_vec<TYPE>(p[v], px);
#pragma unroll
                for (int e = 0; e < T::VPX; ++e) {
                    const int j = v * T::VPX + e;
#pragma unroll
                    for (int k = 0; k < HO; ++k) {
                        const int t = j - k - T::D;
                        if (t == 0) o[k] = mul4p(gr[0], px[e]);
                        else if (t > 0 && t <= 2 * R) o[k] = fma4p(gr[t], px[e], o[k]);
                    }
                }
            }
            if (i0 + lane < ITEMS) {
                float4* m = strip + row * T::MP + half * HO;
#pragma unroll
                for (int k = 0; k < HO; ++k) m[k] = o[k];
            }
        }
        __syncwarp();
        auto release = [&]() {
            if (lane == 0 && release_count(&released[s]) == ST_THREADS / 32 - 1) {
                released[s] = 0;
                if (tile + NBUF * stride < Q.n_tiles) {
                    fence_proxy_async();
                    issue(s, tile + NBUF * stride);
                }
            }
        };
        if (!UNSHARP) release();
        // ---- columns: lane = column x 4 segments of 6 rows ----
        float4 f[PX];
        {
            const float4* p = strip + (seg * PX) * T::MP + col;
#pragma unroll
            for (int j = 0; j < PX + 2 * R; ++j) {
                const float4 q = p[j * T::MP];
#pragma unroll
                for (int k = 0; k < PX; ++k) {
                    const int t = j - k;
                    if (t == 0) f[k] = mul4p(g[0], q);
                    else if (t > 0 && t <= 2 * R) f[k] = fma4p(g[t], q, f[k]);
                }
            }
        }
        const int x = x0 + xb * ST_HB + col, y = y0 + seg * PX;
        if (UNSHARP) {
#pragma unroll
            for (int k = 0; k < PX; ++k)
                f[k] = unsharp_px(staged_px<TYPE>(in, (seg * PX + k + R) * T::PITCH + (xb * ST_HB + col + T::RA) * T::B), f[k], Q.c.contrast, Q.c.threshold);
            __syncwarp();
            release();
        }
        if (x < Q.w) {
#pragma unroll
            for (int k = 0; k < PX; ++k)
                if (y + k < Q.h) store_px(dstp, TYPE, (size_t)(y + k) * Q.w + x, EPI ? epi_apply<TYPE>(Q.epi, f[k]) : f[k]);
        }
    }
}

// k_gauss_march: Blur at large radii (R = 6 .. 9 live taps per side) as a MARCH down the frame (round 2).  A tile of k_gauss_tma
// stages and row-filters TH + 2R rows for TH rows of output -- 42 for 24 at R = 9, three quarters more than it keeps.  Here a
// CTA owns a strip of 64 columns and a chunk of rows and walks down it in bands of 24 source rows: each band is staged once (TMA,
// two buffers), row-filtered once into the warp's private RING of 24 + 2R row slots, and the column pass then emits the 24 rows
// whose window the band completed.  Per output pixel: 2 (2R + 1) taps instead of (2 + 2R / 24)(2R + 1), 57 % of the staged
// bytes, the same tap order as k_gauss_tma (bit-identical frames).  Frame edges: the left / right surround is patched into the
// landed band like k_gauss_tma does; rows above / below the frame are never filtered -- the column pass clamps its row index
// into the ring.
template <int R, int TYPE, int TH_ = 24, int NBUF_ = 2>
struct GmTile {
    static constexpr int TH = TH_, PX = TH_ / 4, NBUF = NBUF_;
    static constexpr int RING = TH + 2 * R;
    static constexpr int MP = ST_HB + 1;
    using G = GtTile<R, TYPE, 2>;               // box width / pitch / alignment of a staged row in the storage type
    static constexpr int BAND_BYTES = (TH * G::PITCH + 127) / 128 * 128;
    static constexpr int RING_BYTES = 8 * RING * MP * 16;
    static constexpr size_t SMEM = NBUF * BAND_BYTES + RING_BYTES + 64;
    static constexpr unsigned TILE_BYTES = (unsigned)(TH * G::PITCH);
};

struct GmParams {
    int w, h, rows_per_chunk;
    ConvParams c;
    EpiSteps epi;
    alignas(64) CUtensorMap map;
};

template <int R, int TYPE, bool EPI, int TH_, int NBUF_>
__global__ void __launch_bounds__(ST_THREADS, 2) k_gauss_march(void* __restrict__ dstp, const __grid_constant__ GmParams Q)
{
    using T = GmTile<R, TYPE, TH_, NBUF_>;
    constexpr int NBUF = NBUF_;
    using G = typename T::G;
    constexpr int PX = T::PX, TH = T::TH, RING = T::RING;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* buf0 = smem_raw;
    unsigned char* buf1 = smem_raw + (NBUF - 1) * T::BAND_BYTES; // (== buf0 with one staging buffer)
    float4* rings = (float4*)(smem_raw + NBUF * T::BAND_BYTES);
    uint64_t* full = (uint64_t*)(smem_raw + NBUF * T::BAND_BYTES + T::RING_BYTES); // [2]
    int* released = (int*)(full + 2);                                           // [2]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int x0 = blockIdx.x * ST_TW;
    const int cy0 = blockIdx.y * Q.rows_per_chunk, cy1 = min(cy0 + Q.rows_per_chunk, Q.h);
    // band b holds source rows [s0 + TH b, s0 + TH (b + 1)); after it the output rows [s0 - R + TH b, s0 - R + TH (b + 1)) are complete
    const int s0 = cy0 - R;
    const int nb = (cy1 - cy0 + 2 * R + TH - 1) / TH;

    auto issue = [&](int s, int band) {
        mbar_expect_tx(&full[s], T::TILE_BYTES);
        tma_load_tile(s ? buf1 : buf0, &Q.map, (x0 - G::RA) * (G::B / G::ELEM), s0 + band * TH, &full[s]);
    };
    if (threadIdx.x == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        released[0] = released[1] = 0;
        mbar_init_fence();
        issue(0, 0);
        if (NBUF > 1 && nb > 1) issue(1, 1);
    }
    __syncthreads();

    // taps centred on R; the row pass carries the load conversion's scale for integer storage (as in k_gauss_tma)
    float g[2 * R + 1], gr[2 * R + 1];
    const float in_scale = TYPE == TC_U8 ? 1.0f / 255.0f : (TYPE == TC_U16 ? 1.0f / 65535.0f : 1.0f);
#pragma unroll
    for (int t = 0; t <= 2 * R; ++t) {
        const int sft = t - R + Q.c.R;
        g[t] = (sft >= 0 && sft <= 2 * Q.c.R) ? Q.c.g[sft] : 0.0f;
        gr[t] = g[t] * in_scale;
    }
    float4* ring = rings + warp * (RING * T::MP);
    const int col = lane % ST_HB, seg = lane / ST_HB;
    const bool xborder = x0 - G::RA < 0 || x0 - G::RA + G::BW > Q.w;
    const int x = x0 + warp * ST_HB + col;

    for (int band = 0; band < nb; ++band) {
        const int s = NBUF > 1 ? (band & 1) : 0;
        unsigned char* in = s ? buf1 : buf0;
        const int r_first = s0 + band * TH; // first source row of the band
        mbar_wait(&full[s], (NBUF > 1 ? (band >> 1) : band) & 1);
        if (xborder) {
            // columns outside the frame <- the clamped in-frame column of the same row (TMA filled them with zeros)
            for (int c = threadIdx.x; c < TH * G::BW; c += ST_THREADS) {
                const int r = c / G::BW, cc = c - r * G::BW;
                const int sx = x0 - G::RA + cc;
                const int cx = min(max(sx, 0), Q.w - 1);
                if (cx != sx) {
                    const unsigned char* from = in + r * G::PITCH + (cx - (x0 - G::RA)) * G::B;
                    unsigned char* to = in + r * G::PITCH + cc * G::B;
                    if (G::B == 16) *reinterpret_cast<uint4*>(to) = *reinterpret_cast<const uint4*>(from);
                    else if (G::B == 8) *reinterpret_cast<uint2*>(to) = *reinterpret_cast<const uint2*>(from);
                    else *reinterpret_cast<unsigned*>(to) = *reinterpret_cast<const unsigned*>(from);
                }
            }
            __syncthreads();
        }
        __syncwarp();
        // ---- rows: lane = source row of the band (rows outside the frame are skipped), 8 outputs from 8 + 2R reads
        {
            const int row = min(lane, TH - 1);
            const int r = r_first + row;
            const uint4* p = reinterpret_cast<const uint4*>(in + row * G::PITCH + (warp * ST_HB) * G::B);
            constexpr int NVI = (ST_HB + 2 * R + G::D + G::VPX - 1) / G::VPX; // 16-byte reads per row item
            float4 o[ST_HB];
#pragma unroll
            for (int v = 0; v < NVI; ++v) {
                float4 px[G::VPX];
                unpack_vec<TYPE>(p[v], px);
#pragma unroll
                for (int e = 0; e < G::VPX; ++e) {
                    const int j = v * G::VPX + e;
#pragma unroll
                    for (int k = 0; k < ST_HB; ++k) {
                        const int t = j - k - G::D;
                        if (t == 0) o[k] = mul4p(gr[0], px[e]);
                        else if (t > 0 && t <= 2 * R) o[k] = fma4p(gr[t], px[e], o[k]);
                    }
                }
            }
            if (lane < TH && r >= 0 && r < Q.h) {
                float4* m = ring + (((band * TH + row) % RING) * T::MP);
#pragma unroll
                for (int k = 0; k < ST_HB; ++k) m[k] = o[k];
            }
        }
        __syncwarp();
        if (lane == 0 && release_count(&released[s]) == ST_THREADS / 32 - 1) {
            released[s] = 0;
            if (band + NBUF < nb) {
                fence_proxy_async();
                issue(s, band + NBUF);
            }
        }
        // ---- columns: lane = column x 4 segments of 6 rows of the 24 output rows this band completed
        const int o_first = r_first - R;                 // first completed output row
        const int y = o_first + seg * PX;
        // ring slot of source row q: (q - s0) mod RING; the window of output row y + k is source rows y + k - R .. y + k + R
        float4 f[PX];
        const bool yedge = o_first - R < 0 || o_first + TH + R > Q.h;
        if (!yedge) {
            int slot = (y - R - s0) % RING;
            const float4* base = ring + col;
#pragma unroll
            for (int j = 0; j < PX + 2 * R; ++j) {
                const float4 q = base[slot * T::MP];
                slot = slot + 1 == RING ? 0 : slot + 1;
#pragma unroll
                for (int k = 0; k < PX; ++k) {
                    const int t = j - k;
                    if (t == 0) f[k] = mul4p(g[0], q);
                    else if (t > 0 && t <= 2 * R) f[k] = fma4p(g[t], q, f[k]);
                }
            }
        } else {
            // a window that reaches over the top / bottom of the frame: the row index clamps (the clamped rows are still in the ring)
            const float4* base = ring + col;
#pragma unroll
            for (int j = 0; j < PX + 2 * R; ++j) {
                const int qrow = min(max(y - R + j, 0), Q.h - 1);
                const float4 q = base[((qrow - s0) % RING) * T::MP];
#pragma unroll
                for (int k = 0; k < PX; ++k) {
                    const int t = j - k;
                    if (t == 0) f[k] = mul4p(g[0], q);
                    else if (t > 0 && t <= 2 * R) f[k] = fma4p(g[t], q, f[k]);
                }
            }
        }
        if (x < Q.w) {
#pragma unroll
            for (int k = 0; k < PX; ++k)
                if (y + k >= cy0 && y + k < cy1) store_px(dstp, TYPE, (size_t)(y + k) * Q.w + x, EPI ? epi_apply<TYPE>(Q.epi, f[k]) : f[k]);
        }
        __syncwarp(); // the ring is rewritten by the next band's row pass
    }
}

template <int R, int TYPE, bool EPI>
static int launch_gauss_march(const tc_image* dst, const tc_image* src, const ConvParams& P, cudaStream_t st, const EpiSteps* epi)
{
    // 32-row bands (every lane of the row pass has a row, 8 output rows per lane of the column pass) with ONE staging buffer
    // where that fits two CTAs per SM, which then cover each other's copies; else 24-row bands with two buffers
    constexpr bool TALL = GmTile<R, TYPE, 32, 1>::SMEM <= 113 * 1024;
    constexpr bool TWO = GmTile<R, TYPE, 24, 2>::SMEM <= 113 * 1024;
    using T = GmTile<R, TYPE, TALL ? 32 : 24, (TALL || !TWO) ? 1 : 2>;
    using G = typename T::G;
    static_assert(T::SMEM <= 113 * 1024, "two CTAs per SM");
    GmParams Q;
    Q.w = dst->w;
    Q.h = dst->h;
    Q.c = P;
    if (EPI) Q.epi = *epi;
    else Q.epi.n = 0;
    // chunks of rows: two waves of two CTAs per SM where the frame allows, at least 8 bands per chunk (2R warm-up rows per chunk)
    const int strips = (dst->w + ST_TW - 1) / ST_TW;
    int chunks = std::max(1, 4 * num_sms() / strips);
    chunks = std::min(chunks, std::max(1, dst->h / (8 * T::TH)));
    Q.rows_per_chunk = ((dst->h + chunks - 1) / chunks + T::TH - 1) / T::TH * T::TH;
    const unsigned long long pitch = (unsigned long long)src->w * G::B;
    int e = tensor_map_2d(src->data, G::ELEM, pitch / G::ELEM, (unsigned long long)src->h, pitch, (unsigned)(G::BW * G::B / G::ELEM), (unsigned)T::TH, &Q.map);
    if (e != TC_OK) return e;
    e = ensure_dynamic_smem((const void*)k_gauss_march<R, TYPE, EPI, T::TH, T::NBUF>, T::SMEM);
    if (e != TC_OK) return e;
    const dim3 grid(strips, (dst->h + Q.rows_per_chunk - 1) / Q.rows_per_chunk);
    k_gauss_march<R, TYPE, EPI, T::TH, T::NBUF><<<grid, ST_THREADS, T::SMEM, st>>>(dst->data, Q);
    count_launch();
    return check_cuda(cudaGetLastError(), "k_gauss_march launch");
}
static bool gauss_march_on()
{
    static const bool on = [] {
        const char* e = getenv("TOUCAN_CONV_MARCH");
        return !(e && *e == '0');
    }();
    return on;
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


static bool gt_use_tma()
{
    static const bool on = [] {
        const char* e = getenv("TOUCAN_CONV_TMA");
        return !(e && *e == '0');
    }();
    return on;
}

template <int R, int TYPE, bool UNSHARP, bool EPI = false>
static int launch_gauss_tma(const tc_image* dst, const tc_image* src, const ConvParams& P, cudaStream_t st, const EpiSteps* epi = nullptr)
{
    // two staging buffers per CTA unless that leaves one CTA (8 warps) per SM: then one buffer and two CTAs, which cover each
    // other's copy latency (float frames from R = 6 up)
    constexpr int NBUF = GtTile<R, TYPE, 2>::SMEM <= 113 * 1024 ? 2 : (GtTile<R, TYPE, 1>::SMEM <= 113 * 1024 ? 1 : 2);
    using T = GtTile<R, TYPE, NBUF>;
    static_assert(T::SMEM <= 227 * 1024, "tile does not fit in shared memory");
    constexpr int OCC = T::SMEM <= 75 * 1024 ? 3 : (T::SMEM <= 113 * 1024 ? 2 : 1);
    GtParams Q;
    Q.w = dst->w;
    Q.h = dst->h;
    Q.tiles_x = (dst->w + ST_TW - 1) / ST_TW;
    Q.n_tiles = Q.tiles_x * ((dst->h + T::TH - 1) / T::TH);
    Q.c = P;
    if (EPI) Q.epi = *epi;
    else Q.epi.n = 0;
    const unsigned long long pitch = (unsigned long long)src->w * T::B;
    int e = tensor_map_2d(src->data, T::ELEM, pitch / T::ELEM, (unsigned long long)src->h, pitch, (unsigned)(T::BW * T::B / T::ELEM), (unsigned)T::IH, &Q.map);
    if (e != TC_OK) return e;
    e = ensure_dynamic_smem((const void*)k_gauss_tma<R, TYPE, UNSHARP, NBUF, OCC, EPI>, T::SMEM);
    if (e != TC_OK) return e;
    const int grid = min(Q.n_tiles, num_sms() * OCC);
    k_gauss_tma<R, TYPE, UNSHARP, NBUF, OCC, EPI><<<grid, ST_THREADS, T::SMEM, st>>>(dst->data, Q);
    count_launch();
    return check_cuda(cudaGetLastError(), "k_gauss_tma launch");
}

template <int R, bool UNSHARP, bool EPI = false>
static int launch_gauss_tma_t(const tc_image* dst, const tc_image* src, const ConvParams& P, cudaStream_t st, const EpiSteps* epi = nullptr)
{
    if constexpr (!UNSHARP && R >= 6) {
        if (gauss_march_on() && dst->h >= 64) {
            switch (src->type) {
            case TC_U8: return launch_gauss_march<R, TC_U8, EPI>(dst, src, P, st, epi);
            case TC_U16: return launch_gauss_march<R, TC_U16, EPI>(dst, src, P, st, epi);
            case TC_F16: return launch_gauss_march<R, TC_F16, EPI>(dst, src, P, st, epi);
            default: return launch_gauss_march<R, TC_F32, EPI>(dst, src, P, st, epi);
            }
        }
    }
    switch (src->type) {
    case TC_U8: return launch_gauss_tma<R, TC_U8, UNSHARP, EPI>(dst, src, P, st, epi);
    case TC_U16: return launch_gauss_tma<R, TC_U16, UNSHARP, EPI>(dst, src, P, st, epi);
    case TC_F16: return launch_gauss_tma<R, TC_F16, UNSHARP, EPI>(dst, src, P, st, epi);
    default: return launch_gauss_tma<R, TC_F32, UNSHARP, EPI>(dst, src, P, st, epi);
    }
}
template <bool UNSHARP>
static int launch_gauss_epi(const tc_image* dst, const tc_image* src, const ConvParams& P, cudaStream_t st, const EpiSteps* epi)
{
    if (P.R <= 2) return launch_gauss_tma_t<2, UNSHARP, true>(dst, src, P, st, epi);
    if (P.R <= 4) return launch_gauss_tma_t<4, UNSHARP, true>(dst, src, P, st, epi);
    if (P.R <= 6) return launch_gauss_tma_t<6, UNSHARP, true>(dst, src, P, st, epi);
    if (P.R <= 9) return launch_gauss_tma_t<9, UNSHARP, true>(dst, src, P, st, epi);
    return launch_gauss_tma_t<12, UNSHARP, true>(dst, src, P, st, epi);
}

// the TMA kernel's preconditions: same-sized, same-typed frames of at least one tile, 16-byte aligned rows
static bool gauss_tma_ok(const tc_image* dst, const tc_image* src)
{
    if (!gt_use_tma() || !tma_available()) return false;
    if (src->w != dst->w || src->h != dst->h || src->type != dst->type || dst->w < ST_TW || dst->h < 24) return false;
    const size_t pitch = (size_t)src->w * 4 * tc_type_size(src->type);
    return !(pitch & 15) && !((uintptr_t)src->data & 15);
}
// (Half / 8-bit frames at R >= 6 were a few percent faster in round 1's typed-load kernel -- u8 r = 20: 0.099 vs 0.114 ms at 4K --
// because the TMA kernel converts a staged pixel once per warp that reads it.  They take the TMA kernel all the same: a node chain
// fused into its epilogue must give the very bits of the node-by-node path, so both have to run the same convolution.)

template <bool UNSHARP>
static int launch_conv_t(const tc_image* dst, const tc_image* src, const ConvParams& P, cudaStream_t st)
{
    if (P.R <= 12 && gauss_tma_ok(dst, src)) {
        if (P.R <= 2) return launch_gauss_tma_t<2, UNSHARP>(dst, src, P, st);
        if (P.R <= 4) return launch_gauss_tma_t<4, UNSHARP>(dst, src, P, st);
        if (P.R <= 6) return launch_gauss_tma_t<6, UNSHARP>(dst, src, P, st);
        if (P.R <= 9) return launch_gauss_tma_t<9, UNSHARP>(dst, src, P, st);
        return launch_gauss_tma_t<12, UNSHARP>(dst, src, P, st);
    }
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

int gauss_with_epilogue(const tc_image* dst, const tc_image* src, const tc_chain_op& head, const EpiSteps& epi, tc_stream s, int* fused)
{
    *fused = 0;
    const bool unsharp = TC_CHAIN_UNSHARP == head.kind;
    if (unsharp && strcmp(head.name_a, "gaussian") != 0) return TC_OK; // (the per-node call reports it)
    ConvParams P;
    if (build_params(head.value, &P) != TC_OK || P.R > 12 || !gauss_tma_ok(dst, src)) return TC_OK;
    P.contrast = unsharp ? head.color[0] : 0.f;
    P.threshold = unsharp ? head.color[1] : 0.f;
    *fused = 1;
    return unsharp ? launch_gauss_epi<true>(dst, src, P, (cudaStream_t)s, &epi) : launch_gauss_epi<false>(dst, src, P, (cudaStream_t)s, &epi);
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
