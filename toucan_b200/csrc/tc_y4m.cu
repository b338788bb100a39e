// tc_y4m.cu -- y4m output conversion on the device.
//
// Replaces, for `toucan-render -y4m FORMAT`, the host-side chain of
// bin/toucan-render/App.cpp:340-485:  ImageBufAlgo::paste into a u8/u16 RGB(A) buffer, then
// libswscale sws_scale_frame(RGB24|RGBA|RGB48 -> YUV422P|YUV444P|YUVA444P|YUV444P16,
// SWS_FAST_BILINEAR), then the planes written back to back.  libswscale is not in the
// reference tree (FFmpeg 8.0, cmake/SuperBuild/BuildFFmpeg.cmake:390); its x86-64 behaviour for
// these four conversions is restated here and in oracle/y4m.py, and pinned bit-exactly against
// the real libswscale 9.1.100 (tests/golden/make_y4m_golden.py).  What that behaviour is:
//   * packed RGB -> 14-bit Y, and U/V from the SUM of each horizontal pixel pair (swscale halves
//     chroma horizontally on RGB input whenever SWS_FAST_BILINEAR is set, even for 4:4:4 output);
//   * a horizontal 2-tap filter per plane.  For the 8-bit outputs the x increment carries
//     swscale's fast-bilinear fudge ((srcW-2)<<16)/(dstW-2) - 20, so "same size" is in fact a
//     very slight stretch whose phase drifts by 20/65536 pixel per column; for 4:4:4 the
//     half-width chroma is stretched back to full width by the same filter;
//   * 15-bit (19-bit for 16-bit output) intermediates, rounded to the output depth.
// One kernel: a CTA converts one segment of one row -- pixels are loaded once (coalesced),
// turned into integer Y/U/V/A samples in shared memory, then filtered from there and written
// as planar bytes.  HBM-bound: reads the frame once, writes 2-6 bytes per pixel.
#include <algorithm>
#include <cstdlib>
#include <map>
#include <mutex>
#include <tuple>
#include <type_traits>
#include <vector>

#include "tc_common.cuh"

namespace tc {

// swscale's default (ITU-R BT.601, limited range) RGB -> YUV table, RGB2YUV_SHIFT = 15
constexpr int RY = 8414, GY = 16519, BY = 3208;
constexpr int RU = -4865, GU = -9528, BU = 14392;
constexpr int RV = 14392, GV = -12061, BV = -2332;

constexpr int Y4M_SEG = 1024;   // output pixels of one row per CTA
constexpr int Y4M_HALO = 64;    // extra source pixels a segment may need (phase drift + taps)
constexpr int Y4M_THREADS = 256;

struct Tap { int pos; int w; }; // taps at pos and pos + 1; w = w0 | w1 << 16

struct Y4mParams {
    const Tap* luma;   // one per output column
    const Tap* chroma; // one per chroma output column
    int cw_src;        // half-width chroma samples per row
    int cw_dst;        // chroma plane width
};

// ---- host: swscale's initFilter() for the bilinear branch ----------------------------
static int64_t rounded_div(int64_t a, int64_t b) { return a >= 0 ? (a + (b >> 1)) / b : (a - (b >> 1)) / b; }

static bool sws_filter(std::vector<Tap>& out, int64_t xInc, int srcW, int dstW)
{
    out.resize((size_t)dstW);
    const int srcPos = 128, dstPos = 128; // get_local_pos() of both sides for these formats
    const int64_t one = 1 << 14;
    if (llabs(xInc - 0x10000) < 10 && srcPos == dstPos) {
        for (int i = 0; i < dstW; ++i) out[(size_t)i] = Tap { i, (int)one };
        return true;
    }
    const int64_t fone = (int64_t)1 << 54;
    const int fs = 2, fsz = 4; // two live taps, storage aligned to 4 (x86)
    std::vector<int64_t> f((size_t)dstW * fsz, 0);
    std::vector<int> pos((size_t)dstW);
    int64_t x = (((int64_t)dstPos * xInc) >> 8) - (((int64_t)srcPos * 0x8000) >> 7);
    for (int i = 0; i < dstW; ++i) {
        int xx = (int)((x - ((fs - 1) << 15) + (1 << 15)) >> 16);
        pos[(size_t)i] = xx;
        for (int j = 0; j < fs; ++j) {
            int64_t c = fone - llabs((int64_t)xx * 65536 - x) * (fone >> 16);
            f[(size_t)i * fsz + j] = c < 0 ? 0 : c;
            ++xx;
        }
        x += xInc;
    }
    // near-zero leading coefficients are dropped (position moves right), monotone positions kept
    const double cut = 0.002 * (double)fone;
    for (int i = dstW - 1; i >= 0; --i) {
        int64_t* fi = &f[(size_t)i * fsz];
        double acc = 0.0;
        for (int j = 0; j < fs; ++j) {
            acc += (double)llabs(fi[0]);
            if (acc > cut) break;
            if (i < dstW - 1 && pos[(size_t)i] >= pos[(size_t)i + 1]) break;
            fi[0] = fi[1];
            fi[1] = 0;
            ++pos[(size_t)i];
        }
    }
    for (int i = 0; i < dstW; ++i) {
        int64_t* fi = &f[(size_t)i * fsz];
        int& p = pos[(size_t)i];
        if (p < 0) {
            for (int j = 1; j < fsz; ++j) {
                const int left = j + p > 0 ? j + p : 0;
                fi[left] += fi[j];
                fi[j] = 0;
            }
            p = 0;
        }
        if (p + fsz > srcW) {
            const int shift = p + (fsz - srcW < 0 ? fsz - srcW : 0);
            int64_t acc = 0;
            for (int j = fsz - 1; j >= 0; --j)
                if (p + j >= srcW) { acc += fi[j]; fi[j] = 0; }
            for (int j = fsz - 1; j >= 0; --j) fi[j] = j < shift ? 0 : fi[j - shift];
            p -= shift;
            fi[srcW - 1 - p] += acc;
        }
        int64_t sum = 0;
        for (int j = 0; j < fsz; ++j) sum += fi[j];
        sum = (sum + one / 2) / one;
        if (!sum) sum = 1;
        int64_t error = 0;
        int w[fsz];
        for (int j = 0; j < fsz; ++j) {
            const int64_t v = fi[j] + error;
            const int64_t iv = rounded_div(v, sum);
            w[j] = (int)iv;
            error = v - iv * sum;
        }
        // at most two adjacent live taps: fold the storage offset into the position
        int j0 = 0;
        while (j0 < fsz - 1 && w[j0] == 0) ++j0;
        for (int j = j0 + 2; j < fsz; ++j)
            if (w[j]) return false;
        const int w1 = j0 + 1 < fsz ? w[j0 + 1] : 0;
        if (w[j0] < 0 || w[j0] > 0xffff || w1 < 0 || w1 > 0xffff) return false;
        out[(size_t)i] = Tap { p + j0, w[j0] | (w1 << 16) };
    }
    return true;
}

struct Y4mTables {
    Tap* dev = nullptr; // luma taps then chroma taps
    int cw_src = 0, cw_dst = 0;
};

static int y4m_tables(int format, int w, Y4mTables& out)
{
    static std::mutex mutex;
    static std::map<std::tuple<int, int, int>, Y4mTables> cache; // (device, format, width)
    int device = 0;
    TC_CUDA(cudaGetDevice(&device));
    std::lock_guard<std::mutex> lock(mutex);
    const auto key = std::make_tuple(device, format, w);
    const auto it = cache.find(key);
    if (it != cache.end()) {
        out = it->second;
        return TC_OK;
    }
    const int cw_src = (w + 1) / 2;
    const int cw_dst = format == TC_Y4M_422 ? cw_src : w;
    int64_t lx, cx;
    if (format == TC_Y4M_444P16) {
        lx = (((int64_t)w << 16) + (w >> 1)) / w;
        cx = (((int64_t)cw_src << 16) + (cw_dst >> 1)) / cw_dst;
    } else {
        // swscale's fast-bilinear increments when the MMXEXT scaler is not used (never for RGB input)
        lx = ((int64_t)(w - 2) << 16) / (w - 2) - 20;
        cx = ((int64_t)(cw_src - 2) << 16) / (cw_dst - 2) - 20;
    }
    std::vector<Tap> luma, chroma;
    if (!sws_filter(luma, lx, w, w) || !sws_filter(chroma, cx, cw_src, cw_dst))
        return fail(TC_ERR_UNSUPPORTED, "tc_rgb_to_y4m: width %d gives a filter with more than two live taps", w);
    // every segment's taps must fall inside the window the kernel stages
    for (int s = 0; s < w; s += Y4M_SEG) {
        const int e = s + Y4M_SEG < w ? s + Y4M_SEG : w;
        int lo = luma[(size_t)s].pos, hi = luma[(size_t)e - 1].pos + 1;
        const int cs = format == TC_Y4M_422 ? s / 2 : s, ce = format == TC_Y4M_422 ? (e + 1) / 2 : e;
        lo = std::min(lo, 2 * chroma[(size_t)cs].pos);
        hi = std::max(hi, 2 * (chroma[(size_t)ce - 1].pos + 1) + 1);
        lo &= ~1;
        if (hi - lo + 1 > Y4M_SEG + Y4M_HALO) return fail(TC_ERR_UNSUPPORTED, "tc_rgb_to_y4m: width %d drifts past the staging window", w);
    }
    Y4mTables t;
    t.cw_src = cw_src;
    t.cw_dst = cw_dst;
    TC_CUDA(cudaMalloc((void**)&t.dev, (luma.size() + chroma.size()) * sizeof(Tap)));
    TC_CUDA(cudaMemcpy(t.dev, luma.data(), luma.size() * sizeof(Tap), cudaMemcpyHostToDevice));
    TC_CUDA(cudaMemcpy(t.dev + luma.size(), chroma.data(), chroma.size() * sizeof(Tap), cudaMemcpyHostToDevice));
    cache[key] = t;
    out = t;
    return TC_OK;
}

// ---- device --------------------------------------------------------------------------
template <bool P16>
__device__ __forceinline__ unsigned filt(const unsigned short* s, Tap t, int base, int last)
{
    const int a = t.pos - base;
    const int b = min(a + 1, last);
    const unsigned v = (unsigned)s[a] * (unsigned)(t.w & 0xffff) + (unsigned)s[b] * ((unsigned)t.w >> 16);
    if (P16) {
        const unsigned v19 = min(v >> 11, (1u << 19) - 1);
        return min((v19 + 4) >> 3, 65535u);
    }
    const unsigned v15 = min(v >> 13, (1u << 15) - 1);
    return min((v15 + 64) >> 7, 255u);
}

template <int FORMAT, int TYPE>
__global__ void __launch_bounds__(Y4M_THREADS) k_rgb_to_y4m(void* __restrict__ dst, const Img src, const Y4mParams P)
{
    constexpr bool P16 = FORMAT == TC_Y4M_444P16;
    constexpr bool ALPHA = FORMAT == TC_Y4M_444ALPHA;
    constexpr bool C422 = FORMAT == TC_Y4M_422;
    constexpr int N = Y4M_SEG + Y4M_HALO;
    __shared__ unsigned short sY[N];
    __shared__ unsigned short sU[N / 2], sV[N / 2];
    __shared__ unsigned short sA[ALPHA ? N : 2];

    const int y = blockIdx.y;
    const int w = src.w;
    const int s = blockIdx.x * Y4M_SEG, e = min(s + Y4M_SEG, w);
    const int cs = C422 ? s / 2 : s, ce = C422 ? (e + 1) / 2 : e;
    // source window of this segment: luma taps and (pairs of) chroma taps, from an even pixel
    int lo = min(P.luma[s].pos, 2 * P.chroma[cs].pos) & ~1;
    int hi = max(P.luma[e - 1].pos + 1, 2 * (P.chroma[ce - 1].pos + 1) + 1);
    hi = min(hi, (w - 1) | 1);
    const int pairs = (hi - lo + 2) / 2;
    const size_t row = (size_t)y * w;

    // stage A: packed RGBA (any storage type) -> quantised RGB(A) -> integer Y / pair-sum U, V
    // The source's storage type is a template parameter and all of the thread's pixel pairs are loaded before the first
    // is converted: with a run-time type switch per load each conversion waited for its own load (four to six exposed
    // latencies per thread).
    constexpr int PAIR_ITERS = (N / 2 + Y4M_THREADS - 1) / Y4M_THREADS;
    float4 q0[PAIR_ITERS], q1[PAIR_ITERS];
#pragma unroll
    for (int i = 0; i < PAIR_ITERS; ++i) {
        const int k = threadIdx.x + i * Y4M_THREADS;
        if (k < pairs) {
            const int x0 = lo + 2 * k;
            // a row of odd width ends in half a pair; its partner reads as the last pixel again
            const int x1 = min(x0 + 1, w - 1);
            q0[i] = load_px(src.p, TYPE, row + x0);
            q1[i] = load_px(src.p, TYPE, row + x1);
        }
    }
#pragma unroll
    for (int i = 0; i < PAIR_ITERS; ++i) {
        const int k = threadIdx.x + i * Y4M_THREADS;
        if (k >= pairs) continue;
        const float4 p0 = q0[i], p1 = q1[i];
        if (P16) {
            const unsigned r0 = f_to_u16(p0.x), g0 = f_to_u16(p0.y), b0 = f_to_u16(p0.z);
            const unsigned r1 = f_to_u16(p1.x), g1 = f_to_u16(p1.y), b1 = f_to_u16(p1.z);
            sY[2 * k] = (unsigned short)((RY * r0 + GY * g0 + BY * b0 + (0x2001u << 14)) >> 15);
            sY[2 * k + 1] = (unsigned short)((RY * r1 + GY * g1 + BY * b1 + (0x2001u << 14)) >> 15);
            const int r = (int)((r0 + r1 + 1) >> 1), g = (int)((g0 + g1 + 1) >> 1), b = (int)((b0 + b1 + 1) >> 1);
            sU[k] = (unsigned short)((unsigned)(RU * r + GU * g + BU * b + (int)(0x10001u << 14)) >> 15);
            sV[k] = (unsigned short)((unsigned)(RV * r + GV * g + BV * b + (int)(0x10001u << 14)) >> 15);
        } else {
            const unsigned r0 = f_to_u8(p0.x), g0 = f_to_u8(p0.y), b0 = f_to_u8(p0.z);
            const unsigned r1 = f_to_u8(p1.x), g1 = f_to_u8(p1.y), b1 = f_to_u8(p1.z);
            sY[2 * k] = (unsigned short)((RY * r0 + GY * g0 + BY * b0 + (32u << 14) + (1u << 8)) >> 9);
            sY[2 * k + 1] = (unsigned short)((RY * r1 + GY * g1 + BY * b1 + (32u << 14) + (1u << 8)) >> 9);
            const int r = (int)(r0 + r1), g = (int)(g0 + g1), b = (int)(b0 + b1);
            sU[k] = (unsigned short)((RU * r + GU * g + BU * b + (256 << 15) + (1 << 9)) >> 10);
            sV[k] = (unsigned short)((RV * r + GV * g + BV * b + (256 << 15) + (1 << 9)) >> 10);
            if (ALPHA) {
                const unsigned a0 = f_to_u8(p0.w), a1 = f_to_u8(p1.w);
                sA[2 * k] = (unsigned short)((a0 << 6) | (a0 >> 2));
                sA[2 * k + 1] = (unsigned short)((a1 << 6) | (a1 >> 2));
            }
        }
    }
    __syncthreads();

    // stage B: horizontal taps from shared memory, planar stores
    using Out = typename std::conditional<P16, unsigned short, unsigned char>::type;
    Out* const planeY = reinterpret_cast<Out*>(dst);
    Out* const planeU = planeY + (size_t)w * src.h;
    Out* const planeV = planeU + (size_t)P.cw_dst * src.h;
    Out* const planeA = planeV + (size_t)P.cw_dst * src.h;
    const int lastY = min(hi, w - 1) - lo;
    const int lastC = (min(hi, w - 1) - lo) / 2;
    for (int x = s + threadIdx.x; x < e; x += Y4M_THREADS) {
        const Tap t = P.luma[x];
        planeY[row + x] = (Out)filt<P16>(sY, t, lo, lastY);
        if (ALPHA) planeA[row + x] = (Out)filt<P16>(sA, t, lo, lastY);
    }
    const size_t crow = (size_t)y * P.cw_dst;
    for (int x = cs + threadIdx.x; x < ce; x += Y4M_THREADS) {
        const Tap t = P.chroma[x];
        planeU[crow + x] = (Out)filt<P16>(sU, t, lo / 2, lastC);
        planeV[crow + x] = (Out)filt<P16>(sV, t, lo / 2, lastC);
    }
}

} // namespace tc

using namespace tc;

extern "C" {

size_t tc_y4m_frame_bytes(int format, int w, int h)
{
    if (w <= 0 || h <= 0) return 0;
    const size_t cw = (size_t)(w + 1) / 2, n = (size_t)w * h;
    switch (format) {
    case TC_Y4M_422: return n + 2 * cw * h;
    case TC_Y4M_444: return 3 * n;
    case TC_Y4M_444ALPHA: return 4 * n;
    case TC_Y4M_444P16: return 6 * n;
    default: return 0;
    }
}

int tc_rgb_to_y4m(void* dst_dev, const tc_image* src, int format, tc_stream s)
{
    int v;
    if ((v = validate(src, "tc_rgb_to_y4m src")) != TC_OK) return v;
    if (!dst_dev) return fail(TC_ERR_INVALID, "tc_rgb_to_y4m: null destination");
    if (format < TC_Y4M_422 || format > TC_Y4M_444P16) return fail(TC_ERR_INVALID, "tc_rgb_to_y4m: format %d", format);
    // swscale's fast-bilinear increment divides by (width - 2) of the half-width chroma plane, and
    // an odd width makes it read one pixel past the row: neither case is defined in the reference
    if (src->w < 8 || (src->w & 1)) return fail(TC_ERR_UNSUPPORTED, "tc_rgb_to_y4m: width %d (even widths >= 8)", src->w);
    if (src->h > 65535) return fail(TC_ERR_UNSUPPORTED, "tc_rgb_to_y4m: height %d", src->h);
    Y4mTables t;
    if ((v = y4m_tables(format, src->w, t)) != TC_OK) return v;
    const Y4mParams P { t.dev, t.dev + src->w, t.cw_src, t.cw_dst };
    const dim3 grid((unsigned)((src->w + Y4M_SEG - 1) / Y4M_SEG), (unsigned)src->h);
    cudaStream_t st = (cudaStream_t)s;
#define TC_Y4M_LAUNCH(F)                                                                                  \
    switch (src->type) {                                                                                  \
    case TC_U8: k_rgb_to_y4m<F, TC_U8><<<grid, Y4M_THREADS, 0, st>>>(dst_dev, view(src), P); break;       \
    case TC_U16: k_rgb_to_y4m<F, TC_U16><<<grid, Y4M_THREADS, 0, st>>>(dst_dev, view(src), P); break;     \
    case TC_F16: k_rgb_to_y4m<F, TC_F16><<<grid, Y4M_THREADS, 0, st>>>(dst_dev, view(src), P); break;     \
    default: k_rgb_to_y4m<F, TC_F32><<<grid, Y4M_THREADS, 0, st>>>(dst_dev, view(src), P); break;         \
    }
    switch (format) {
    case TC_Y4M_422: TC_Y4M_LAUNCH(TC_Y4M_422) break;
    case TC_Y4M_444: TC_Y4M_LAUNCH(TC_Y4M_444) break;
    case TC_Y4M_444ALPHA: TC_Y4M_LAUNCH(TC_Y4M_444ALPHA) break;
    default: TC_Y4M_LAUNCH(TC_Y4M_444P16) break;
    }
#undef TC_Y4M_LAUNCH
    count_launch();
    return check_cuda(cudaGetLastError(), "k_rgb_to_y4m launch");
}

} // extern "C"
