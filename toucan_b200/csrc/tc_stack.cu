// tc_stack.cu -- cross-node fusion of the hot chain of a multi-track timeline:
//   per track:  [toucan:Blur] -> [toucan:Dissolve with the neighbouring clip] -> CompNode(premult + over)
// evaluated for ALL tracks in one pass over the output tile, accumulator in registers.
// Compulsory HBM traffic = every referenced source frame read once (+ halo, served by L2)
// and the final frame written once, instead of one read+write of a full frame per graph
// node (lib/toucanRender/ImageGraph.cpp:165-198, Comp.cpp:29-84, FilterPlugin.cpp:188-202,
// TransitionPlugin.cpp:242-264).  float32 RGBA only (the working format of the 8K config).
//
// Kernel shape (k_stack_f32<R>): one CTA of 256 threads per 64 x 24 output tile, three
// CTAs resident per SM.  Per blurred layer:
//   stage   (64+2R) x (24+2R) source pixels -> shared memory, coordinates clamped to the
//           frame (OIIO convolve clamps); a Dissolve whose two inputs carry the same blur
//           is mixed HERE, before the blur (convolution is linear, so blur(a)*iv + blur(b)*v
//           = blur(a*iv + b*v) to float rounding) -- one blur instead of two;
//   rows    horizontal taps: each thread owns 8 adjacent pixels of one row and slides the
//           2R+1 taps over 8+2R shared-memory reads (2 reads per output instead of 2R+1);
//   columns vertical taps: each thread owns 6 consecutive rows of one column (6+2R reads),
//           then blends the finished layer into its 6 accumulator pixels.
// Row strides of both shared tiles are odd multiples of 16 bytes so the row-per-lane
// accesses of the horizontal pass are bank-conflict free.
//
// Tolerance path (<= 1e-5 abs), FMA contraction allowed.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "tc_tile.cuh"
#include "tc_tma.cuh"

namespace tc {

constexpr int ST_MAX_R = 12;      // live taps per side (radius parameter up to ~26)
constexpr int ST_MAX_KERNELS = 4; // distinct blur kernels per launch
constexpr int ST_MAX_LAYERS = 32;
constexpr int ST_MAX_UNITS = ST_MAX_LAYERS;

struct StKernel {
    int R;
    float scale;
    float g[2 * ST_MAX_R + 1]; // already multiplied by sqrt(scale)
};
// One step of the per-tile pipeline = one layer.  PLAIN: a blurred source.  MIX: a blurred
// Dissolve whose two inputs carry the same blur -- both tiles are staged, mixed in place in
// shared memory (a*iv + b*v) and blurred ONCE (convolution is linear).  DIRECT: an unblurred
// layer read straight from global memory.
enum { ST_MIX = 1, ST_DIRECT = 4 };
struct StUnit {
    const float4* src;
    const float4* src2; // MIX / DIRECT Dissolve: the second input (or null)
    float w, w2;        // Dissolve weights (iv, v)
    short kernel;
    short flags;
    int next_staged;    // index of the next staged unit (prefetch target), or -1
    short map, map2;    // k_stack_tma: tensor maps of src / src2
};
struct StParams {
    int w, h, n_units, first_staged;
    int tiles_x, tiles_y, band; // CTA order: bands of `band` tile rows, column-major inside a band
    float4 background;
    StKernel k[ST_MAX_KERNELS];
    StUnit unit[ST_MAX_UNITS];
};

// Dissolve: add(mul(from, iv), mul(to, v)) (float storage: nothing is quantised)
__device__ __forceinline__ float4 mix4(float4 a, float4 b, float iv, float v)
{
    return make_float4(__fadd_rn(__fmul_rn(a.x, iv), __fmul_rn(b.x, v)), __fadd_rn(__fmul_rn(a.y, iv), __fmul_rn(b.y, v)),
                       __fadd_rn(__fmul_rn(a.z, iv), __fmul_rn(b.z, v)), __fadd_rn(__fmul_rn(a.w, iv), __fmul_rn(b.w, v)));
}

__device__ __forceinline__ void cp_async16(float4* smem_dst, const float4* gsrc)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// Asynchronous stage of one source tile (with halo, coordinates clamped to the frame as
// OIIO's convolve does) into shared memory: thread = one staged column, STAGE_ROWS rows per
// pass, no registers held while the copies are in flight.
template <int R, int TH>
__device__ __forceinline__ void tile_prefetch(const float4* __restrict__ src, int w, int h, int x0, int y0, float4* in)
{
    using T = StTile<R, TH>;
    const int tid = threadIdx.x;
    if (tid < T::STAGE_ROWS * T::IW) {
        const int ry = tid / T::IW, ix = tid - ry * T::IW;
        const float4* p = src + min(max(x0 - R + ix, 0), w - 1);
        float4* d = in + ix;
#pragma unroll
        for (int u = 0; u < T::STAGE_N; ++u) {
            const int iy = ry + u * T::STAGE_ROWS;
            if (iy < T::IH) cp_async16(d + iy * T::IWP, p + (size_t)min(max(y0 - R + iy, 0), h - 1) * w);
        }
    }
    cp_async_commit();
}

template <int ST_PX>
__device__ __forceinline__ void tile_load(const float4* __restrict__ src, int w, int h, int x, int y, float4 out[ST_PX])
{
    x = min(x, w - 1);
#pragma unroll
    for (int k = 0; k < ST_PX; ++k) out[k] = __ldg(src + (size_t)min(y + k, h - 1) * w + x);
}

// Two staging buffers + mid, 2 CTAs per SM at radius 10; tiles stream in (cp.async) while the
// previous layer is filtered.
template <int R, int TH, int NBUF, int OCC>
__global__ void __launch_bounds__(ST_THREADS, OCC) k_stack_f32(float4* __restrict__ dst, const __grid_constant__ StParams P)
{
    static_assert(NBUF == 2, "the pipeline uses two staging buffers");
    using T = StTile<R, TH, NBUF>;
    constexpr int ST_PX = T::PX;
    extern __shared__ float4 smem[];
    float4* mid = smem + NBUF * T::IN_ELEMS;
    // CTA -> tile: the CTAs resident at one time (2-3 per SM, ~300-450) walk every source at about the same
    // pace, so a halo row or column shared by two of them is read from HBM once and from L2 the second time
    // -- if both are resident together.  In row-major order a wave is 2-4 full rows of tiles and the halo
    // rows towards the next wave have left L2 (20 sources x 300 tiles x 37 KB = 220 MB) before it starts;
    // bands of `band` tile rows walked column by column keep most vertical neighbours in the same wave.
    const int per_band = P.band * P.tiles_x;
    const int band = blockIdx.x / per_band, within = blockIdx.x - band * per_band;
    const int bh = min(P.band, P.tiles_y - band * P.band);
    const int tx = within / bh, ty = band * P.band + (within - tx * bh);
    const int x0 = tx * ST_TW, y0 = ty * TH;
    const int col = threadIdx.x % ST_TW, seg = threadIdx.x / ST_TW;
    const int x = x0 + col, y = y0 + seg * ST_PX;

    float4 acc[ST_PX];
#pragma unroll
    for (int k = 0; k < ST_PX; ++k) acc[k] = P.background;

    // Buffer roles: the unit being filtered has its tile `a` in buffer `slot` and, for a Dissolve,
    // tile `b` in the other one (mixed on the fly by the row pass).  A plain layer leaves the other
    // buffer free, so the next unit's `a` streams into it during the whole filter; after the row
    // pass both buffers are free and whatever the next unit still needs streams in during the
    // column pass and the blend.  The next unit then sees a / b in (slot^1, slot).
    int slot = 0;
    if (P.first_staged >= 0) {
        const StUnit& N = P.unit[P.first_staged];
        tile_prefetch<R, TH>(N.src, P.w, P.h, x0, y0, smem);
        if (N.flags & ST_MIX) tile_prefetch<R, TH>(N.src2, P.w, P.h, x0, y0, smem + T::IN_ELEMS);
    }

    for (int u = 0; u < P.n_units; ++u) {
        const StUnit& U = P.unit[u];
        float4 f[ST_PX];
        if (U.flags & ST_DIRECT) {
            tile_load<ST_PX>(U.src, P.w, P.h, x, y, f);
            if (U.src2) {
                float4 fb[ST_PX];
                tile_load<ST_PX>(U.src2, P.w, P.h, x, y, fb);
#pragma unroll
                for (int k = 0; k < ST_PX; ++k) f[k] = mix4(f[k], fb[k], U.w, U.w2);
            }
        } else {
            const StKernel& K = P.k[U.kernel];
            // taps centred on R (zero padded when this kernel is narrower than the template)
            float g[2 * R + 1];
#pragma unroll
            for (int t = 0; t <= 2 * R; ++t) {
                const int s = t - R + K.R;
                g[t] = (s >= 0 && s <= 2 * K.R) ? K.g[s] : 0.0f;
            }
            cp_async_wait_all();
            __syncthreads(); // this unit's tile(s) have landed; `mid` is free again
            float4* in = smem + slot * T::IN_ELEMS;
            float4* other = smem + (slot ^ 1) * T::IN_ELEMS;
            const StUnit* N = U.next_staged >= 0 ? &P.unit[U.next_staged] : nullptr;
            const bool mix = (U.flags & ST_MIX) != 0;
            if (mix) {
                tile_rows_mix<R, TH>(in, other, U.w, U.w2, mid, g);
            } else {
                if (N) tile_prefetch<R, TH>(N->src, P.w, P.h, x0, y0, other);
                tile_rows<R, TH, true>(in, mid, g);
            }
            __syncthreads(); // row pass done: `mid` is complete, the staging buffers are free
            if (N && mix) tile_prefetch<R, TH>(N->src, P.w, P.h, x0, y0, other);
            if (N && (N->flags & ST_MIX)) tile_prefetch<R, TH>(N->src2, P.w, P.h, x0, y0, in);
            slot ^= 1;
            tile_cols<R, TH>(mid, g, f);
        }
        // CompNode: premult(fg) then over(fg, acc); tolerance path, so fused multiply-adds
#pragma unroll
        for (int k = 0; k < ST_PX; ++k) {
            float4 p = f[k];
            const float a = p.w;
            const float pm = a != 1.0f ? a : 1.0f;
            p.x *= pm;
            p.y *= pm;
            p.z *= pm;
            const float oma = 1.0f - fminf(fmaxf(a, 0.0f), 1.0f);
            acc[k] = fma4p(oma, acc[k], p);
        }
    }
    if (x < P.w) {
#pragma unroll
        for (int k = 0; k < ST_PX; ++k)
            if (y + k < P.h) __stcs(dst + (size_t)(y + k) * P.w + x, acc[k]);
    }
}


// =====================================================================================================
// k_stack_tma: the same chain with the source tiles moved by the TMA engine and no CTA-wide barrier.
//
//   * One CUtensorMap per source frame (2-D: 2W x H elements of 8 bytes, so a pixel is two elements and the
//     box may be IWP = 64 + 2R + 1 pixels wide -- the odd row pitch the row pass wants -- without exceeding
//     the 256-element box limit).  ONE cp.async.bulk.tensor.2d moves a (64+2R+1) x (24+2R) tile; its completion
//     is counted on an mbarrier ("full").  There is no producer warp (a ninth warp would cost every warp 16
//     registers: the register file is split over four schedulers): a staging buffer is handed back through a
//     shared-memory counter, and the lane whose warp hands it back LAST issues the copy of the tile after next
//     into it -- nobody ever blocks on "empty".  The 8 warps never touch global memory for staged layers.
//   * TMA fills out-of-frame cells with zeros, OIIO's convolve clamps: the ~3 % of tiles that touch the frame
//     border patch those cells from the clamped in-frame cell of the same tile after it lands (the only place a
//     consumer-wide named barrier is used).
//   * The row-pass result is warp-private: warp w filters the 8-column strip w of the tile (lane = source row),
//     writes its own 9-pixel-pitch strip of `mid`, and runs the column pass on the same strip (lane = column x
//     4 row segments of 6), so row -> column needs __syncwarp only.  Warps drift apart and the shared-memory
//     heavy row pass of one overlaps the FMA heavy column pass of another.
template <int R, int TH>
struct StTma {
    static constexpr int PX = 6;                       // rows per lane in the column pass (4 segments x 6 = TH)
    static constexpr int IW = ST_TW + 2 * R, IH = TH + 2 * R, IWP = IW | 1;
    static constexpr int IN_BYTES = ((IH * IWP * 16) + 127) / 128 * 128;
    static constexpr int MP = ST_HB + 1;               // strip pitch (pixels): odd, row-per-lane stores conflict-free
    static constexpr int MID_BYTES = 8 * IH * MP * 16; // 8 warps
    static constexpr int BAR_BYTES = 64; // full[2], (unused) empty[2], released[2]
    static constexpr size_t SMEM = 2 * IN_BYTES + MID_BYTES + BAR_BYTES;
    static_assert(TH == 4 * PX, "column pass: 4 segments of 6 rows");
};
constexpr int ST_MAX_MAPS = 2 * ST_MAX_LAYERS;
constexpr int ST_TMA_THREADS = ST_THREADS;

struct StTmaParams {
    StParams p;
    int n_tiles;                    // staged tiles per CTA, in consumption order
    short tile_map[ST_MAX_MAPS];    // tile -> tensor map
    alignas(64) CUtensorMap maps[ST_MAX_MAPS];
};

__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;\n" ::"n"(ST_THREADS) : "memory"); }

// cells of a landed tile that lie outside the frame <- the clamped in-frame cell of the same tile
template <int R, int TH>
__device__ __forceinline__ void tile_clamp_patch(float4* in, int w, int h, int x0, int y0)
{
    using T = StTma<R, TH>;
    for (int i = threadIdx.x; i < T::IH * T::IW; i += ST_THREADS) {
        const int r = i / T::IW, c = i - r * T::IW;
        const int sx = x0 - R + c, sy = y0 - R + r;
        const int cx = min(max(sx, 0), w - 1), cy = min(max(sy, 0), h - 1);
        if (cx != sx || cy != sy) in[r * T::IWP + c] = in[(cy - (y0 - R)) * T::IWP + (cx - (x0 - R))];
    }
}

template <int R, int TH, bool MIX>
__device__ __forceinline__ void strip_rows(const float4* ina, const float4* inb, float wa, float wb, float4* strip, int xb, int lane,
                                           const float (&g)[2 * R + 1])
{
    using T = StTma<R, TH>;
    for (int row = lane; row < T::IH; row += 32) {
        const int base = row * T::IWP + xb * ST_HB;
        float4* m = strip + row * T::MP;
        float4 o[ST_HB];
#pragma unroll
        for (int j = 0; j < ST_HB + 2 * R; ++j) {
            float4 q = ina[base + j];
            if (MIX) q = fma4p(wa, q, mul4p(wb, inb[base + j]));
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

template <int R, int TH>
__device__ __forceinline__ void strip_cols(const float4* strip, int lane, const float (&g)[2 * R + 1], float4 (&out)[StTma<R, TH>::PX])
{
    using T = StTma<R, TH>;
    const int col = lane % ST_HB, seg = lane / ST_HB;
    const float4* p = strip + (seg * T::PX) * T::MP + col;
#pragma unroll
    for (int j = 0; j < T::PX + 2 * R; ++j) {
        const float4 q = p[j * T::MP];
#pragma unroll
        for (int k = 0; k < T::PX; ++k) {
            const int t = j - k;
            if (t == 0) out[k] = mul4p(g[0], q);
            else if (t > 0 && t <= 2 * R) out[k] = fma4p(g[t], q, out[k]);
        }
    }
}

template <int R, int TH, int OCC>
__global__ void __launch_bounds__(ST_TMA_THREADS, OCC) k_stack_tma(float4* __restrict__ dst, const __grid_constant__ StTmaParams Q)
{
    using T = StTma<R, TH>;
    constexpr int PX = T::PX;
    const StParams& P = Q.p;
    extern __shared__ __align__(128) unsigned char smem_raw[]; // (no pointer arithmetic through integers: keeps LDS / STS)
    unsigned char* base = smem_raw;
    float4* buf0 = (float4*)base;
    float4* buf1 = (float4*)(base + T::IN_BYTES);
    float4* mid = (float4*)(base + 2 * T::IN_BYTES);
    uint64_t* full = (uint64_t*)(base + 2 * T::IN_BYTES + T::MID_BYTES); // [2]
    uint64_t* empty = full + 2;                                          // [2]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int per_band = P.band * P.tiles_x;
    const int band = blockIdx.x / per_band, within = blockIdx.x - band * per_band;
    const int bh = min(P.band, P.tiles_y - band * P.band);
    const int tx = within / bh, ty = band * P.band + (within - tx * bh);
    const int x0 = tx * ST_TW, y0 = ty * TH;

    int* released = (int*)(empty + 2); // [2] warps that have handed the buffer back
    (void)empty;
    constexpr unsigned TILE_BYTES = (unsigned)(T::IH * T::IWP * 16);
    if (threadIdx.x == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        released[0] = released[1] = 0;
        mbar_init_fence();
        // the first two tiles
        for (int t = 0; t < 2 && t < Q.n_tiles; ++t) {
            mbar_expect_tx(&full[t], TILE_BYTES);
            tma_load_tile(t ? buf1 : buf0, &Q.maps[Q.tile_map[t]], 2 * (x0 - R), y0 - R, &full[t]);
        }
    }
    __syncthreads();
    // hand staging buffer s (it held tile `tile`) back; the last warp to do so starts the copy of tile + 2 into it
    auto release = [&](int s, int tile) {
        if (release_count(&released[s]) == ST_THREADS / 32 - 1) {
            released[s] = 0;
            if (tile + 2 < Q.n_tiles) {
                // the warps' generic-proxy reads (and border patches) of this buffer are ordered before the copy's async-proxy writes
                fence_proxy_async();
                mbar_expect_tx(&full[s], TILE_BYTES);
                tma_load_tile(s ? buf1 : buf0, &Q.maps[Q.tile_map[tile + 2]], 2 * (x0 - R), y0 - R, &full[s]);
            }
        }
    };

    // ===== consumers =====
    const int xb = warp;                                 // this warp's 8-column strip
    float4* strip = mid + warp * (T::IH * T::MP);
    const int col = lane % ST_HB, seg = lane / ST_HB;
    const int x = x0 + xb * ST_HB + col, y = y0 + seg * PX;
    const bool border = x0 - R < 0 || y0 - R < 0 || x0 + ST_TW + R > P.w || y0 + TH + R > P.h;

    float4 acc[PX];
#pragma unroll
    for (int k = 0; k < PX; ++k) acc[k] = P.background;

    int seq = 0;
    for (int u = 0; u < P.n_units; ++u) {
        const StUnit& U = P.unit[u];
        float4 f[PX];
        if (U.flags & ST_DIRECT) {
            tile_load<PX>(U.src, P.w, P.h, x, y, f);
            if (U.src2) {
                float4 fb[PX];
                tile_load<PX>(U.src2, P.w, P.h, x, y, fb);
#pragma unroll
                for (int k = 0; k < PX; ++k) f[k] = mix4(f[k], fb[k], U.w, U.w2);
            }
        } else {
            const StKernel& K = P.k[U.kernel];
            float g[2 * R + 1];
#pragma unroll
            for (int t = 0; t <= 2 * R; ++t) {
                const int sft = t - R + K.R;
                g[t] = (sft >= 0 && sft <= 2 * K.R) ? K.g[sft] : 0.0f;
            }
            const bool mix = (U.flags & ST_MIX) != 0;
            const int s = seq & 1;
            float4* ina = s ? buf1 : buf0;
            float4* inb = s ? buf0 : buf1;
            mbar_wait(&full[s], (seq >> 1) & 1);
            if (mix) mbar_wait(&full[s ^ 1], ((seq + 1) >> 1) & 1);
            if (border) {
                tile_clamp_patch<R, TH>(ina, P.w, P.h, x0, y0);
                if (mix) tile_clamp_patch<R, TH>(inb, P.w, P.h, x0, y0);
                consumer_sync();
            }
            __syncwarp(); // the previous layer's column pass has left this warp's strip
            if (mix) strip_rows<R, TH, true>(ina, inb, U.w, U.w2, strip, xb, lane, g);
            else strip_rows<R, TH, false>(ina, ina, 1.0f, 0.0f, strip, xb, lane, g);
            __syncwarp();
            if (lane == 0) {
                release(s, seq); // this warp is done with the staged tile(s)
                if (mix) release(s ^ 1, seq + 1);
            }
            seq += mix ? 2 : 1;
            strip_cols<R, TH>(strip, lane, g, f);
        }
#pragma unroll
        for (int k = 0; k < PX; ++k) {
            float4 p = f[k];
            const float a = p.w;
            const float pm = a != 1.0f ? a : 1.0f;
            p.x *= pm;
            p.y *= pm;
            p.z *= pm;
            const float oma = 1.0f - fminf(fmaxf(a, 0.0f), 1.0f);
            acc[k] = fma4p(oma, acc[k], p);
        }
    }
    if (x < P.w) {
#pragma unroll
        for (int k = 0; k < PX; ++k)
            if (y + k < P.h) __stcs(dst + (size_t)(y + k) * P.w + x, acc[k]);
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
    // the two 1-D passes each carry sqrt(scale): their product is the normalised 2-D kernel
    const float root = sqrtf(K->scale);
    memset(K->g, 0, sizeof(K->g));
    for (int i = -live; i <= live; ++i) K->g[i + live] = g[r + i] * root;
    return TC_OK;
}

// CTAs per SM the shared-memory footprint allows (227 KB usable, 1 KB reserved per CTA)
template <int R, int TH, int NBUF>
constexpr int st_occ() { return StTile<R, TH, NBUF>::SMEM <= 74 * 1024 ? 3 : (StTile<R, TH, NBUF>::SMEM <= 112 * 1024 ? 2 : 1); }

static int st_band()
{
    static const int band = [] {
        const char* e = getenv("TOUCAN_STACK_BAND");
        const int v = e ? atoi(e) : 0;
        return v > 0 ? v : 8;
    }();
    return band;
}

template <int R, int TH, int NBUF, int OCC = st_occ<R, TH, NBUF>()>
static int launch_stack(const tc_image* dst, StParams& P, cudaStream_t st)
{
    using T = StTile<R, TH, NBUF>;
    static_assert(T::SMEM <= 227 * 1024, "tile does not fit in shared memory");
    int st_attr = ensure_dynamic_smem((const void*)k_stack_f32<R, TH, NBUF, OCC>, T::SMEM);
    if (st_attr != TC_OK) return st_attr;
    P.tiles_x = (dst->w + ST_TW - 1) / ST_TW;
    P.tiles_y = (dst->h + TH - 1) / TH;
    P.band = min(st_band(), P.tiles_y);
    k_stack_f32<R, TH, NBUF, OCC><<<P.tiles_x * P.tiles_y, ST_THREADS, T::SMEM, st>>>((float4*)dst->data, P);
    count_launch();
    return check_cuda(cudaGetLastError(), "k_stack_f32 launch");
}

static bool st_use_tma()
{
    static const bool on = [] {
        const char* e = getenv("TOUCAN_STACK_TMA");
        return !(e && *e == '0');
    }();
    return on;
}

template <int R, int TH>
static int launch_stack_tma(const tc_image* dst, StParams& P, cudaStream_t st)
{
    using T = StTma<R, TH>;
    static_assert(T::SMEM <= 227 * 1024, "tile does not fit in shared memory");
    constexpr int OCC = T::SMEM <= 113 * 1024 ? 2 : 1;
    StTmaParams Q;
    int n_maps = 0;
    for (int u = 0; u < P.n_units; ++u) {
        StUnit& U = P.unit[u];
        U.map = U.map2 = -1;
        if (U.flags & ST_DIRECT) continue;
        if (n_maps + 2 > ST_MAX_MAPS) return fail(TC_ERR_UNSUPPORTED, "tc_stack_f32: too many staged sources");
        int e = tensor_map_2d(U.src, 8, 2ull * P.w, (unsigned long long)P.h, 16ull * P.w, 2u * T::IWP, (unsigned)T::IH, &Q.maps[n_maps]);
        if (e != TC_OK) return e;
        Q.tile_map[n_maps] = (short)n_maps;
        U.map = (short)n_maps++;
        if (U.flags & ST_MIX) {
            e = tensor_map_2d(U.src2, 8, 2ull * P.w, (unsigned long long)P.h, 16ull * P.w, 2u * T::IWP, (unsigned)T::IH, &Q.maps[n_maps]);
            if (e != TC_OK) return e;
            Q.tile_map[n_maps] = (short)n_maps;
            U.map2 = (short)n_maps++;
        }
    }
    int st_attr = ensure_dynamic_smem((const void*)k_stack_tma<R, TH, OCC>, T::SMEM);
    if (st_attr != TC_OK) return st_attr;
    P.tiles_x = (dst->w + ST_TW - 1) / ST_TW;
    P.tiles_y = (dst->h + TH - 1) / TH;
    P.band = min(st_band(), P.tiles_y);
    Q.p = P;
    Q.n_tiles = n_maps;
    k_stack_tma<R, TH, OCC><<<P.tiles_x * P.tiles_y, ST_TMA_THREADS, T::SMEM, st>>>((float4*)dst->data, Q);
    count_launch();
    return check_cuda(cudaGetLastError(), "k_stack_tma launch");
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
    P.background = make_float4(background[0], background[1], background[2], background[3]);
    float radii[ST_MAX_KERNELS];
    int nk = 0, maxR = 1;
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
        if (P.k[nk].R > maxR) maxR = P.k[nk].R;
        radii[nk] = radius;
        *idx = nk++;
        return TC_OK;
    };
    int nu = 0, last_staged = -1;
    P.first_staged = -1;
    for (int l = 0; l < n_layers; ++l) {
        if (!layers[l].src_a) return fail(TC_ERR_INVALID, "tc_stack_f32: layer %d has no source", l);
        int ka = -1, kb = -1;
        if ((v = kernel_of(layers[l].radius_a, &ka)) != TC_OK) return v;
        if (layers[l].src_b && (v = kernel_of(layers[l].radius_b, &kb)) != TC_OK) return v;
        if (layers[l].src_b && ka != kb)
            return fail(TC_ERR_UNSUPPORTED, "tc_stack_f32: layer %d dissolves inputs with different blurs (%g, %g)", l,
                        layers[l].radius_a, layers[l].radius_b);
        StUnit& U = P.unit[nu];
        U.src = (const float4*)layers[l].src_a;
        U.src2 = (const float4*)layers[l].src_b;
        U.w = (float)(1.0 - layers[l].value);
        U.w2 = (float)layers[l].value;
        U.kernel = (short)ka;
        U.next_staged = -1;
        if (ka < 0) {
            U.flags = ST_DIRECT;
        } else {
            U.flags = layers[l].src_b ? ST_MIX : 0;
            if (last_staged >= 0) {
                P.unit[last_staged].next_staged = nu;
            } else {
                P.first_staged = nu;
            }
            last_staged = nu;
        }
        ++nu;
    }
    P.n_units = nu;
    cudaStream_t st = (cudaStream_t)s;
    // the TMA kernel: frames at least one tile large whose rows are 16-byte multiples (always) and whose base is 16-byte aligned
    bool tma = st_use_tma() && tma_available() && dst->w >= ST_TW && dst->h >= 24;
    for (int u = 0; u < nu && tma; ++u)
        if (((uintptr_t)P.unit[u].src | (uintptr_t)P.unit[u].src2) & 15) tma = false;
    if (tma) {
        if (maxR <= 2) return launch_stack_tma<2, 24>(dst, P, st);
        if (maxR <= 4) return launch_stack_tma<4, 24>(dst, P, st);
        if (maxR <= 6) return launch_stack_tma<6, 24>(dst, P, st);
        if (maxR <= 9) return launch_stack_tma<9, 24>(dst, P, st);
        return launch_stack_tma<12, 24>(dst, P, st);
    }
    if (maxR <= 2) return launch_stack<2, 24, 2>(dst, P, st);
    if (maxR <= 4) return launch_stack<4, 24, 2>(dst, P, st);
    if (maxR <= 6) return launch_stack<6, 24, 2>(dst, P, st);
    if (maxR <= 9) return launch_stack<9, 24, 2>(dst, P, st);
    return launch_stack<12, 24, 2>(dst, P, st);
}
