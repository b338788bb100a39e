// tc_draw.cu -- the Draw effects' pixel work: toucan:Box and toucan:Line.
//
// plugins/DrawPlugin.cpp:198-233 (Box) and :306-341 (Line) copy the source into the output and
// call ImageBufAlgo::render_box / render_line on it.  OIIO draws with "colour over pixel":
//     r[c] = color[c] + r[c] * (1 - alpha),   alpha = color[alpha channel]
// (plain store when alpha == 1), a filled box over ROI [x1, x2] x [y1, y2] clipped to the image,
// an unfilled box as four lines that each skip their first point, a degenerate box as one
// point, and lines with the integer Bresenham walk.  A Bresenham walk is sequential on the CPU;
// here step i's pixel comes from the closed form k_i = max(0, ceil((2*d_minor*i - d_major) /
// (2*d_major))) minor-axis advances before step i, so every step is an independent thread.
// Bit-exact class: unfused multiply/add (-fmad=false), storage-type quantisation on store.
#include "tc_common.cuh"

namespace tc {

struct DrawColor { float c[4]; float one_minus_alpha; int opaque; };

__device__ __forceinline__ float4 draw_over(float4 p, const DrawColor& k)
{
    if (k.opaque) return make_float4(k.c[0], k.c[1], k.c[2], k.c[3]);
    return make_float4(__fadd_rn(k.c[0], __fmul_rn(p.x, k.one_minus_alpha)), __fadd_rn(k.c[1], __fmul_rn(p.y, k.one_minus_alpha)),
                       __fadd_rn(k.c[2], __fmul_rn(p.z, k.one_minus_alpha)), __fadd_rn(k.c[3], __fmul_rn(p.w, k.one_minus_alpha)));
}

// copy (with storage conversion) + filled box in one pass; the box is [bx0, bx1) x [by0, by1), already clipped
__global__ void __launch_bounds__(256) k_copy_box(const Img dst, const Img src, int bx0, int by0, int bx1, int by1, const DrawColor k)
{
    const size_t total = (size_t)dst.w * dst.h;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int y = (int)(i / dst.w), x = (int)(i - (size_t)y * dst.w);
        // ImageBufAlgo::copy into the output's storage type first: the drawn pixel blends with the converted value
        float4 p = quant4(load_black(src, x, y), dst.type);
        if (x >= bx0 && x < bx1 && y >= by0 && y < by1) p = draw_over(p, k);
        store_px(dst.p, dst.type, i, p);
    }
}

// one thread per Bresenham step, in place on dst
__global__ void __launch_bounds__(128) k_line(const Img dst, int x1, int y1, int xinc, int yinc, int dmajor, int dminor, int steep, int first,
                                              const DrawColor k)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x + first;
    if (i > dmajor) return;
    long long num = 2ll * dminor * i - dmajor;
    int adv = 0;
    if (num > 0) adv = (int)((num + 2ll * dmajor - 1) / (2ll * dmajor));
    const int x = steep ? x1 + adv * xinc : x1 + i * xinc;
    const int y = steep ? y1 + i * yinc : y1 + adv * yinc;
    if (x < 0 || y < 0 || x >= dst.w || y >= dst.h) return;
    const size_t at = (size_t)y * dst.w + x;
    store_px(dst.p, dst.type, at, draw_over(load_px(dst.p, dst.type, at), k));
}

static DrawColor draw_color(const float color[4])
{
    DrawColor k;
    for (int c = 0; c < 4; ++c) k.c[c] = color[c];
    k.one_minus_alpha = 1.0f - color[3];
    k.opaque = color[3] == 1.0f;
    return k;
}

static int launch_copy_box(const tc_image* dst, const tc_image* src, int bx0, int by0, int bx1, int by1, const DrawColor& k, cudaStream_t st)
{
    const size_t total = (size_t)dst->w * dst->h;
    size_t blocks = (total + 255) / 256;
    const size_t cap = (size_t)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    k_copy_box<<<(unsigned)blocks, 256, 0, st>>>(view(dst), view(src), bx0, by0, bx1, by1, k);
    count_launch();
    return check_cuda(cudaGetLastError(), "k_copy_box launch");
}

static int launch_line(const tc_image* dst, int x1, int y1, int x2, int y2, const DrawColor& k, bool skip_first, cudaStream_t st)
{
    const int dx = abs(x2 - x1), dy = abs(y2 - y1);
    const int xinc = x1 > x2 ? -1 : 1, yinc = y1 > y2 ? -1 : 1;
    const int steep = dx >= dy ? 0 : 1;
    const int dmajor = steep ? dy : dx, dminor = steep ? dx : dy;
    const int first = skip_first ? 1 : 0;
    const int steps = dmajor + 1 - first;
    if (steps <= 0) return TC_OK;
    k_line<<<(unsigned)((steps + 127) / 128), 128, 0, st>>>(view(dst), x1, y1, xinc, yinc, dmajor, dminor, steep, first, k);
    count_launch();
    return check_cuda(cudaGetLastError(), "k_line launch");
}

} // namespace tc

using namespace tc;

extern "C" {

int tc_draw_box(const tc_image* dst, const tc_image* src, int x1, int y1, int x2, int y2, const float color[4], int fill, tc_stream s)
{
    int v;
    if ((v = validate(dst, "tc_draw_box dst")) != TC_OK) return v;
    if ((v = validate(src, "tc_draw_box src")) != TC_OK) return v;
    if (!color) return fail(TC_ERR_INVALID, "tc_draw_box: null color");
    const DrawColor k = draw_color(color);
    cudaStream_t st = (cudaStream_t)s;
    if (x1 == x2 && y1 == y2) {
        // degenerate box: render_point, one clipped pixel
        return launch_copy_box(dst, src, x1, y1, x1 + 1, y1 + 1, k, st);
    }
    if (fill) {
        // roi_intersection(image, ROI(x1, x2 + 1, y1, y2 + 1)): empty when the corners are not in order
        return launch_copy_box(dst, src, x1 > 0 ? x1 : 0, y1 > 0 ? y1 : 0, x2 + 1 < dst->w ? x2 + 1 : dst->w, y2 + 1 < dst->h ? y2 + 1 : dst->h, k, st);
    }
    if ((v = launch_copy_box(dst, src, 0, 0, 0, 0, k, st)) != TC_OK) return v;
    if ((v = launch_line(dst, x1, y1, x2, y1, k, true, st)) != TC_OK) return v;
    if ((v = launch_line(dst, x2, y1, x2, y2, k, true, st)) != TC_OK) return v;
    if ((v = launch_line(dst, x2, y2, x1, y2, k, true, st)) != TC_OK) return v;
    return launch_line(dst, x1, y2, x1, y1, k, true, st);
}

int tc_draw_line(const tc_image* dst, const tc_image* src, int x1, int y1, int x2, int y2, const float color[4], int skip_first_point, tc_stream s)
{
    int v;
    if ((v = validate(dst, "tc_draw_line dst")) != TC_OK) return v;
    if ((v = validate(src, "tc_draw_line src")) != TC_OK) return v;
    if (!color) return fail(TC_ERR_INVALID, "tc_draw_line: null color");
    const DrawColor k = draw_color(color);
    cudaStream_t st = (cudaStream_t)s;
    if ((v = launch_copy_box(dst, src, 0, 0, 0, 0, k, st)) != TC_OK) return v;
    return launch_line(dst, x1, y1, x2, y2, k, skip_first_point != 0, st);
}

} // extern "C"
