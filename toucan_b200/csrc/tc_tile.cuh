// tc_tile.cuh -- the register-blocked separable-filter tile shared by the fused stack kernel
// (tc_stack.cu) and the per-node Blur / UnsharpMask kernel (tc_conv.cu).
//
// A CTA of 256 threads owns a 64 x TH output tile.  The source tile (+R halo each side) sits in
// shared memory; the horizontal pass gives every thread 8 adjacent pixels of one row (8 + 2R
// shared-memory reads for 8 outputs), the vertical pass TH/4 consecutive rows of one column.
// Row strides are odd multiples of 16 bytes, so the row-per-lane accesses of the horizontal pass
// are bank-conflict free.  Arithmetic is packed fp32x2 (FFMA2): the tile is issue-bound.
#pragma once

#include "tc_common.cuh"

namespace tc {

constexpr int ST_TW = 64, ST_THREADS = 256;
constexpr int ST_HB = 8; // pixels per thread in the horizontal pass

// Packed fp32x2 arithmetic (FFMA2 / FMUL2): one issue slot per two multiply-adds.  The
// kernel is issue-bound, not FMA-pipe-bound, so halving the FMA issue slots is the lever.
__device__ __forceinline__ float4 fma4p(float w, float4 p, float4 a)
{
    const float2 ww = make_float2(w, w);
    const float2 lo = __ffma2_rn(ww, make_float2(p.x, p.y), make_float2(a.x, a.y));
    const float2 hi = __ffma2_rn(ww, make_float2(p.z, p.w), make_float2(a.z, a.w));
    return make_float4(lo.x, lo.y, hi.x, hi.y);
}
__device__ __forceinline__ float4 mul4p(float w, float4 p)
{
    const float2 ww = make_float2(w, w);
    const float2 lo = __fmul2_rn(ww, make_float2(p.x, p.y));
    const float2 hi = __fmul2_rn(ww, make_float2(p.z, p.w));
    return make_float4(lo.x, lo.y, hi.x, hi.y);
}

template <int R, int TH, int NBUF = 2>
struct StTile {
    static constexpr int PX = ST_TW * TH / ST_THREADS; // output pixels per thread: one column segment
    static constexpr int IW = ST_TW + 2 * R;  // staged columns
    static constexpr int IH = TH + 2 * R;     // staged rows
    static constexpr int IWP = IW | 1;        // odd row stride (float4 units)
    static constexpr int MWP = ST_TW | 1;
    static constexpr int IN_ELEMS = IH * IWP;
    static constexpr int MID_ELEMS = IH * MWP;
    static constexpr size_t SMEM = (size_t)(NBUF * IN_ELEMS + MID_ELEMS) * sizeof(float4); // staging buffer(s) + mid
    static constexpr int STAGE_ROWS = ST_THREADS / IW;                 // rows staged per pass
    static constexpr int STAGE_N = (IH + STAGE_ROWS - 1) / STAGE_ROWS; // passes
};

// rows: 8 adjacent outputs per thread, lanes on consecutive rows; FIRST writes mid, else adds
template <int R, int TH, bool FIRST>
__device__ __forceinline__ void tile_rows(const float4* in, float4* mid, const float (&g)[2 * R + 1])
{
    using T = StTile<R, TH>;
    constexpr int N_H = T::IH * (ST_TW / ST_HB);
    for (int item = threadIdx.x; item < N_H; item += ST_THREADS) {
        const int row = item % T::IH, xb = item / T::IH;
        const float4* p = in + row * T::IWP + xb * ST_HB;
        float4* m = mid + row * T::MWP + xb * ST_HB;
        float4 o[ST_HB];
        if (!FIRST) {
#pragma unroll
            for (int k = 0; k < ST_HB; ++k) o[k] = m[k];
        }
#pragma unroll
        for (int j = 0; j < ST_HB + 2 * R; ++j) {
            const float4 q = p[j];
#pragma unroll
            for (int k = 0; k < ST_HB; ++k) {
                const int t = j - k;
                if (FIRST && t == 0) o[k] = mul4p(g[0], q);
                else if (t >= 0 && t <= 2 * R) o[k] = fma4p(g[t], q, o[k]);
            }
        }
#pragma unroll
        for (int k = 0; k < ST_HB; ++k) m[k] = o[k];
    }
}

// rows of the Dissolve of two staged tiles, mixed on the fly: every input is read once from each
// tile (a*wa + b*wb) and fed to the taps -- no separate mixing pass over shared memory
template <int R, int TH>
__device__ __forceinline__ void tile_rows_mix(const float4* ina, const float4* inb, float wa, float wb, float4* mid,
                                              const float (&g)[2 * R + 1])
{
    using T = StTile<R, TH>;
    constexpr int N_H = T::IH * (ST_TW / ST_HB);
    for (int item = threadIdx.x; item < N_H; item += ST_THREADS) {
        const int row = item % T::IH, xb = item / T::IH;
        const int base = row * T::IWP + xb * ST_HB;
        const float4* pa = ina + base;
        const float4* pb = inb + base;
        float4* m = mid + row * T::MWP + xb * ST_HB;
        float4 o[ST_HB];
#pragma unroll
        for (int j = 0; j < ST_HB + 2 * R; ++j) {
            const float4 q = fma4p(wa, pa[j], mul4p(wb, pb[j]));
#pragma unroll
            for (int k = 0; k < ST_HB; ++k) {
                const int t = j - k;
                if (t == 0) o[k] = mul4p(g[0], q);
                else if (t > 0 && t <= 2 * R) o[k] = fma4p(g[t], q, o[k]);
            }
        }
#pragma unroll
        for (int k = 0; k < ST_HB; ++k) m[k] = o[k];
    }
}

// columns: PX consecutive rows per thread, lanes on consecutive columns
template <int R, int TH>
__device__ __forceinline__ void tile_cols(const float4* mid, const float (&g)[2 * R + 1], float4 out[StTile<R, TH>::PX])
{
    using T = StTile<R, TH>;
    const int col = threadIdx.x % ST_TW, seg = threadIdx.x / ST_TW;
    const float4* p = mid + (seg * T::PX) * T::MWP + col;
#pragma unroll
    for (int j = 0; j < T::PX + 2 * R; ++j) {
        const float4 q = p[j * T::MWP];
#pragma unroll
        for (int k = 0; k < T::PX; ++k) {
            const int t = j - k;
            if (t == 0) out[k] = mul4p(g[0], q);
            else if (t > 0 && t <= 2 * R) out[k] = fma4p(g[t], q, out[k]);
        }
    }
}

} // namespace tc
