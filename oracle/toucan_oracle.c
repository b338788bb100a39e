/*
 * toucan_oracle.c -- CPU ORACLE for the toucan per-frame render path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library.  The
 * product path (toucan_b200/) never links, loads or calls it.
 *
 * PARITY: COARSELY PINNED.  The pixel arithmetic of the reference lives in OpenImageIO
 * v3.0.10.0 (cmake/SuperBuild/BuildOpenImageIO.cmake:4) and OpenColorIO v2.4.2
 * (BuildOpenColorIO.cmake:4), neither of which is vendored under /root/reference nor
 * installed here, and the reference's own tests hold no golden pixels
 * (tests/toucanRenderTest/ImageGraphTest.cpp:62-74 only writes PNGs).  Each function
 * below restates the published OIIO algorithm that the cited toucan call site invokes.
 * What pins the restatement to the reference's output: the film strips the reference
 * rendered itself (the images/ PNG strips), matched per frame at thumbnail scale to <= 1/255 mean
 * by tests/test_reference_filmstrips.py (every effect of the fixtures except toucan:Text).
 * Below that resolution the items restated from the third-party libraries stay [V] (to
 * verify against a real OIIO build when one is available).  Constants cross-checked
 * here: fast_sinpi (max abs error 0.000918958 measured vs 0.000918954611 documented)
 * and the colour-map knots (OpenCV's copy of the matplotlib tables, linearised);
 * see tests/test_oracle_pins.py.
 *
 * Conventions: every image is interleaved RGBA (4 channels), row-major, top row
 * first, no row padding (PropertySet.cpp:467, Read.cpp:79-84).  All arithmetic is
 * float32 with storage-type conversion at load/store.  Compile with
 * -ffp-contract=off: OIIO is built for baseline x86-64 (no FMA).
 */
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { O_U8 = 0, O_U16 = 1, O_F16 = 2, O_F32 = 3 };

typedef struct {
    void* data;
    int w, h;
    int type;
} oimg;

#define NCH 4

/* bench.py's CPU arms: the OpenMP team size, set explicitly (torchrun exports OMP_NUM_THREADS=1 to its ranks) */
void o_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }
int o_get_threads(void)
{
    int n = 1;
#pragma omp parallel
    {
#pragma omp single
        n = omp_get_num_threads();
    }
    return n;
}

/* ------------------------------------------------------------------------- */
/* 9.1 storage conversion  [V: OIIO fmath.h convert_type / scaled_conversion] */

static inline float o_clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }

static inline uint8_t f2u8(float f)
{
    float s = f * 255.0f;
    s += (s < 0.0f ? -0.5f : 0.5f);
    s = o_clampf(s, 0.0f, 255.0f);
    return (uint8_t)s;
}
static inline uint16_t f2u16(float f)
{
    float s = f * 65535.0f;
    s += (s < 0.0f ? -0.5f : 0.5f);
    s = o_clampf(s, 0.0f, 65535.0f);
    return (uint16_t)s;
}
static inline float u82f(uint8_t v) { return (float)v * (1.0f / 255.0f); }
static inline float u162f(uint16_t v) { return (float)v * (1.0f / 65535.0f); }

static inline float ldc(const oimg* im, int x, int y, int c)
{
    size_t i = ((size_t)y * im->w + x) * NCH + c;
    switch (im->type) {
    case O_U8: return u82f(((const uint8_t*)im->data)[i]);
    case O_U16: return u162f(((const uint16_t*)im->data)[i]);
    case O_F16: return (float)((const _Float16*)im->data)[i];
    default: return ((const float*)im->data)[i];
    }
}
static inline void stc(oimg* im, int x, int y, int c, float v)
{
    size_t i = ((size_t)y * im->w + x) * NCH + c;
    switch (im->type) {
    case O_U8: ((uint8_t*)im->data)[i] = f2u8(v); break;
    case O_U16: ((uint16_t*)im->data)[i] = f2u16(v); break;
    case O_F16: ((_Float16*)im->data)[i] = (_Float16)v; break;
    default: ((float*)im->data)[i] = v; break;
    }
}
static inline void ldpx(const oimg* im, int x, int y, float* p)
{
    for (int c = 0; c < NCH; ++c) p[c] = ldc(im, x, y, c);
}
/* black outside the data window (OIIO WrapBlack / reads outside the data window) */
static inline void ldpx_black(const oimg* im, int x, int y, float* p)
{
    if (x < 0 || y < 0 || x >= im->w || y >= im->h) {
        p[0] = p[1] = p[2] = p[3] = 0.0f;
    } else
        ldpx(im, x, y, p);
}
static inline void ldpx_clamp(const oimg* im, int x, int y, float* p)
{
    x = x < 0 ? 0 : (x >= im->w ? im->w - 1 : x);
    y = y < 0 ? 0 : (y >= im->h ? im->h - 1 : y);
    ldpx(im, x, y, p);
}
static inline void stpx(oimg* im, int x, int y, const float* p)
{
    for (int c = 0; c < NCH; ++c) stc(im, x, y, c, p[c]);
}
/* float -> storage type -> float round trip */
static inline float quant(float v, int type)
{
    switch (type) {
    case O_U8: return u82f(f2u8(v));
    case O_U16: return u162f(f2u16(v));
    case O_F16: return (float)(_Float16)v;
    default: return v;
    }
}

size_t o_type_size(int type)
{
    static const size_t s[4] = { 1, 2, 2, 4 };
    return s[type & 3];
}

/* [V: OIIO TypeDesc::basetype_merge] result type of op(A,B) into an uninitialised dst */
int o_type_merge(int a, int b)
{
    if (a == b) return a;
    size_t sa = o_type_size(a), sb = o_type_size(b);
    if (sa < sb) { int t = a; a = b; b = t; }
    if (a == O_F32) return O_F32;
    if ((a == O_U16 || a == O_F16) && b == O_U8) return a;
    return O_F32;
}

/* OIIO fmath.h lerp: v0*(1-x) + v1*x */
static inline float o_lerp(float a, float b, float x) { return a * (1.0f - x) + b * x; }

/* ------------------------------------------------------------------------- */
/* 9.2 generators */

/* ImageBufAlgo::fill(dst, values)  -- GeneratorPlugin.cpp:296-303 */
void o_fill(oimg* dst, const float* color)
{
#pragma omp parallel for schedule(static)
    for (int y = 0; y < dst->h; ++y)
        for (int x = 0; x < dst->w; ++x) stpx(dst, x, y, color);
}

/* ImageBufAlgo::fill(dst, top, bottom)  -- GeneratorPlugin.cpp:393-407 (vertical gradient)
   [V] v = (y - ybegin) / max(1, height-1); lerp per channel */
void o_fill_tb(oimg* dst, const float* top, const float* bottom)
{
    float hh = (float)(dst->h - 1 > 1 ? dst->h - 1 : 1);
#pragma omp parallel for schedule(static)
    for (int y = 0; y < dst->h; ++y) {
        float v = (float)y / hh;
        float p[NCH];
        for (int c = 0; c < NCH; ++c) p[c] = o_lerp(top[c], bottom[c], v);
        for (int x = 0; x < dst->w; ++x) stpx(dst, x, y, p);
    }
}

/* Horizontal gradient -- GeneratorPlugin.cpp:409-431: a UINT8 image of size
   (width=H_out, height=W_out) gets a vertical ramp, is rotated 270 and copied into the
   output.  rotate270: dst(x', y') = src(W-1-y', x')  => value depends on x' only.
   => out(x, y) = convert(u8q(lerp(c1, c2, x / max(1, W_out-1)))) */
void o_gradient_h(oimg* dst, const float* c1, const float* c2)
{
    float hh = (float)(dst->w - 1 > 1 ? dst->w - 1 : 1);
#pragma omp parallel for schedule(static)
    for (int y = 0; y < dst->h; ++y)
        for (int x = 0; x < dst->w; ++x) {
            float v = (float)x / hh;
            float p[NCH];
            for (int c = 0; c < NCH; ++c) p[c] = quant(o_lerp(c1[c], c2[c], v), O_U8);
            stpx(dst, x, y, p);
        }
}

/* ImageBufAlgo::checker(dst, w, h, 1, c1, c2)  -- GeneratorPlugin.cpp:205-221  [V] */
void o_checker(oimg* dst, int cw, int ch, const float* c1, const float* c2)
{
    if (cw <= 0 || ch <= 0) return; /* reference divides by zero; undefined */
#pragma omp parallel for schedule(static)
    for (int y = 0; y < dst->h; ++y)
        for (int x = 0; x < dst->w; ++x) {
            int v = x / cw + y / ch; /* + z/1 with z = 0 */
            stpx(dst, x, y, (v & 1) ? c2 : c1);
        }
}

/* Bob Jenkins lookup3 final mix  [V: OIIO hash.h bjhash::bjfinal] */
static inline uint32_t rotl32(uint32_t x, int k) { return (x << k) | (x >> (32 - k)); }
static inline uint32_t bjfinal(uint32_t a, uint32_t b, uint32_t c)
{
    c ^= b; c -= rotl32(b, 14);
    a ^= c; a -= rotl32(c, 11);
    b ^= a; b -= rotl32(a, 25);
    c ^= b; c -= rotl32(b, 16);
    a ^= c; a -= rotl32(c, 4);
    b ^= a; b -= rotl32(a, 14);
    c ^= b; c -= rotl32(b, 24);
    return c;
}
static inline float hashrand(int x, int y, int z, int c, int seed)
{
    const uint32_t magic = 0xfffff;
    uint32_t h = bjfinal(bjfinal((uint32_t)x, (uint32_t)y, (uint32_t)z), (uint32_t)c, (uint32_t)seed) & magic;
    return (float)h * (1.0f / (float)(magic + 1));
}

/* Natural log in double built only from IEEE + - * / (the classic fdlibm e_log
   algorithm), so that a CUDA restatement of the same sequence is bit-identical.  The
   reference calls libm log(); this differs from glibc by < 1 ulp of double, which is
   invisible after the narrowing to float in hashnormal. */
static double det_log(double x)
{
    static const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10,
                        Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01,
                        Lg3 = 2.857142874366239149e-01, Lg4 = 2.222219843214978396e-01,
                        Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
                        Lg7 = 1.479819860511658591e-01;
    uint64_t bits;
    memcpy(&bits, &x, 8);
    int32_t hx = (int32_t)(bits >> 32);
    uint32_t lx = (uint32_t)bits;
    int32_t k = 0, i, j;
    /* callers only pass normal positive finite values (r2 in (0,1]) */
    k += (hx >> 20) - 1023;
    hx &= 0x000fffff;
    i = (hx + 0x95f64) & 0x100000;
    bits = ((uint64_t)(uint32_t)(hx | (i ^ 0x3ff00000)) << 32) | lx;
    memcpy(&x, &bits, 8);
    k += (i >> 20);
    double f = x - 1.0, dk = (double)k;
    if ((0x000fffff & (2 + hx)) < 3) {
        if (f == 0.0) {
            if (k == 0) return 0.0;
            return dk * ln2_hi + dk * ln2_lo;
        }
        double R = f * f * (0.5 - 0.33333333333333333 * f);
        if (k == 0) return f - R;
        return dk * ln2_hi - ((R - dk * ln2_lo) - f);
    }
    double s = f / (2.0 + f);
    double z = s * s;
    i = hx - 0x6147a;
    double w = z * z;
    j = 0x6b851 - hx;
    double t1 = w * (Lg2 + w * (Lg4 + w * Lg6));
    double t2 = z * (Lg1 + w * (Lg3 + w * (Lg5 + w * Lg7)));
    i |= j;
    double R = t2 + t1;
    if (i > 0) {
        double hfsq = 0.5 * f * f;
        if (k == 0) return f - (hfsq - s * (hfsq + R));
        return dk * ln2_hi - ((hfsq - (s * (hfsq + R) + dk * ln2_lo)) - f);
    }
    if (k == 0) return f - s * (f - R);
    return dk * ln2_hi - ((s * (f - R) - dk * ln2_lo) - f);
}
double o_det_log(double x) { return det_log(x); }

/* Marsaglia polar method  [V: OIIO imagebufalgo_draw.cpp hashnormal] */
static inline float hashnormal(int x, int y, int z, int c, int seed)
{
    float xr, yr, r2;
    int s = seed - 1;
    do {
        s++;
        xr = (float)(2.0 * (double)hashrand(x, y, z, c, s) - 1.0);
        yr = (float)(2.0 * (double)hashrand(x, y, z, c, s + 139) - 1.0);
        r2 = xr * xr + yr * yr;
    } while (r2 > 1.0f || r2 == 0.0f);
    float M = (float)sqrt(-2.0 * det_log((double)r2) / (double)r2);
    return xr * M;
}

/* ImageBufAlgo::noise(dst, type, A, B, mono, seed) -- GeneratorPlugin.cpp:525-531.
   Adds to the existing (zero-initialised, ImageEffect.cpp:93) pixel values.
   type: 0 = gaussian/normal, 1 = white/uniform, 2 = salt.  */
void o_noise(oimg* dst, int type, float A, float B, int mono, int seed)
{
#pragma omp parallel for schedule(static)
    for (int y = 0; y < dst->h; ++y)
        for (int x = 0; x < dst->w; ++x) {
            float p[NCH];
            ldpx(dst, x, y, p);
            float n = 0.0f;
            for (int c = 0; c < NCH; ++c) {
                if (type == 0) {
                    if (c == 0 || !mono) n = A + B * hashnormal(x, y, 0, c, seed);
                    p[c] = p[c] + n;
                } else if (type == 1) {
                    if (c == 0 || !mono) n = A + (B - A) * hashrand(x, y, 0, c, seed);
                    p[c] = p[c] + n;
                } else {
                    if (c == 0 || !mono) n = hashrand(x, y, 0, c, seed);
                    if (n < B) p[c] = A;
                }
            }
            stpx(dst, x, y, p);
        }
}

/* ------------------------------------------------------------------------- */
/* 9.3 pointwise filters */

/* invert(dst, src, roi ch[0,3)) + copy alpha -- FilterPlugin.cpp:335-361.
   [V] invert == mad(src, -1, 1) */
void o_invert(oimg* dst, const oimg* src)
{
#pragma omp parallel for schedule(static)
    for (int y = 0; y < dst->h; ++y)
        for (int x = 0; x < dst->w; ++x) {
            float p[NCH];
            ldpx_black(src, x, y, p);
            for (int c = 0; c < 3; ++c) p[c] = p[c] * -1.0f + 1.0f;
            stpx(dst, x, y, p);
        }
}

/* pow(dst, src, value) on all four channels -- FilterPlugin.cpp:430-438 */
void o_pow(oimg* dst, const oimg* src, float value)
{
#pragma omp parallel for schedule(static)
    for (int y = 0; y < dst->h; ++y)
        for (int x = 0; x < dst->w; ++x) {
            float p[NCH];
            ldpx_black(src, x, y, p);
            for (int c = 0; c < NCH; ++c) p[c] = powf(p[c], value);
            stpx(dst, x, y, p);
        }
}

static inline float luma709(const float* p) { return (0.2126f * p[0] + 0.7152f * p[1]) + 0.0722f * p[2]; }

/* saturate(dst, src, scale, firstchannel=0) -- FilterPlugin.cpp:507-516  [V] */
void o_saturate(oimg* dst, const oimg* src, float scale)
{
#pragma omp parallel for schedule(static)
    for (int y = 0; y < dst->h; ++y)
        for (int x = 0; x < dst->w; ++x) {
            float p[NCH];
            ldpx_black(src, x, y, p);
            float l = luma709(p);
            for (int c = 0; c < 3; ++c) p[c] = o_lerp(l, p[c], scale);
            stpx(dst, x, y, p);
        }
}

/* Colour-map knot tables [V: OIIO imagebufalgo_pixelmath.cpp color_map]: 17 knots per matplotlib map / turbo in
   LINEAR light (the published sRGB-encoded map at indices 0,16,...,240,255, linearised).  Generated by
   tools/make_colormaps.py; pinned at thumbnail scale by the reference's own render images/Filter.png
   (tests/test_reference_filmstrips.py: 0.24/255 mean with these knots, 34.5/255 with round 1's sRGB-encoded ones). */
#include "colormap_tables.inc"

/* returns number of knots and pointer, or 0 for an unknown name */
int o_color_map_knots(const char* name, const float** knots)
{
#define CM(n, t) if (!strcmp(name, n)) { *knots = t; return (int)(sizeof(t) / sizeof(float) / 3); }
    CM("magma", k_magma) CM("inferno", k_inferno) CM("plasma", k_plasma) CM("viridis", k_viridis)
    CM("turbo", k_turbo) CM("blue-red", k_bluered) CM("red-blue", k_bluered) CM("bluered", k_bluered)
    CM("redblue", k_bluered) CM("spectrum", k_spectrum) CM("heat", k_heat)
#undef CM
    return 0;
}

/* [V: OIIO fmath.h interpolate_linear over a strided span] */
static inline float interp_linear(float x, const float* knots, int n, int c)
{
    x = o_clampf(x, 0.0f, 1.0f);
    int nsegs = n - 1;
    float xs = x * (float)nsegs;
    float fl = floorf(xs);
    int seg = (int)fl;
    float fr = xs - fl;
    int next = seg + 1 < nsegs ? seg + 1 : nsegs;
    return o_lerp(knots[seg * 3 + c], knots[next * 3 + c], fr);
}

/* color_map(dst, src, -1, name) + copy alpha -- FilterPlugin.cpp:271-295.
   An unknown map name makes color_map fail (dst RGB stays zero) and alpha is copied. */
void o_color_map(oimg* dst, const oimg* src, const char* name)
{
    const float* knots = NULL;
    int n = o_color_map_knots(name, &knots);
#pragma omp parallel for schedule(static)
    for (int y = 0; y < dst->h; ++y)
        for (int x = 0; x < dst->w; ++x) {
            float p[NCH], q[NCH];
            ldpx_black(src, x, y, p);
            float l = luma709(p);
            for (int c = 0; c < 3; ++c) q[c] = n ? interp_linear(l, knots, n, c) : 0.0f;
            q[3] = p[3];
            stpx(dst, x, y, q);
        }
}

/* premult(dst, src) -- ColorPlugin.cpp:286, Comp.cpp:41,80  [V] */
void o_premult(oimg* dst, const oimg* src)
{
#pragma omp parallel for schedule(static)
    for (int y = 0; y < dst->h; ++y)
        for (int x = 0; x < dst->w; ++x) {
            float p[NCH];
            ldpx_black(src, x, y, p);
            float a = p[3];
            if (a != 1.0f)
                for (int c = 0; c < 3; ++c) p[c] = p[c] * a;
            stpx(dst, x, y, p);
        }
}

/* unpremult(dst, src) -- ColorPlugin.cpp:324  [V] */
void o_unpremult(oimg* dst, const oimg* src)
{
#pragma omp parallel for schedule(static)
    for (int y = 0; y < dst->h; ++y)
        for (int x = 0; x < dst->w; ++x) {
            float p[NCH];
            ldpx_black(src, x, y, p);
            float a = p[3];
            if (a != 0.0f && a != 1.0f)
                for (int c = 0; c < 3; ++c) p[c] = p[c] / a;
            stpx(dst, x, y, p);
        }
}

/* ------------------------------------------------------------------------- */
/* 9.4 convolution family */

/* [V: OIIO fmath.h fast_exp2/fast_exp]  no FMA (baseline x86-64) */
static inline float fast_exp2(float xval)
{
    float x = o_clampf(xval, -126.0f, 126.0f);
    int m = (int)x;
    x -= (float)m;
    x = 1.0f - (1.0f - x);
    float r = 1.33336498402e-3f;
    r = x * r + 9.810352697968e-3f;
    r = x * r + 5.551834031939e-2f;
    r = x * r + 0.2401793301105f;
    r = x * r + 0.693144857883f;
    r = x * r + 1.0f;
    uint32_t bits;
    memcpy(&bits, &r, 4);
    bits += ((uint32_t)m << 23);
    memcpy(&r, &bits, 4);
    return r;
}
static inline float fast_exp(float x) { return fast_exp2(x * (float)(1.0 / M_LN2)); }
float o_fast_exp(float x) { return fast_exp(x); }

/* make_kernel("gaussian", width, width) [V: OIIO make_kernel + FilterGaussian2D]:
   size n = ceil(width) rounded up to odd; value g(2x/w) g(2y/h), g(u) = u<1 ? fast_exp(-2u^2) : 0;
   normalised to sum 1.  The 2-D kernel is returned in k2d (n*n floats).  Returns n. */
int o_make_gauss_kernel(float width, float* k2d, int cap)
{
    int n = (int)ceilf(width);
    if (n < 1) n = 1;
    if (!(n & 1)) ++n;
    if (!k2d) return n;
    if (n * n > cap) return -n;
    float radinv = 2.0f / width;
    int r = n / 2;
    float sum = 0.0f;
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) {
            float xx = fabsf((float)(i - r) * radinv), yy = fabsf((float)(j - r) * radinv);
            float gx = xx < 1.0f ? fast_exp(-2.0f * (xx * xx)) : 0.0f;
            float gy = yy < 1.0f ? fast_exp(-2.0f * (yy * yy)) : 0.0f;
            float v = gx * gy;
            k2d[j * n + i] = v;
            sum += v;
        }
    if (sum != 0.0f)
        for (int i = 0; i < n * n; ++i) k2d[i] = k2d[i] / sum;
    return n;
}

/* convolve(dst, src, K, normalize=true) [V]: clamp edges, float accumulate j outer / i
   inner over the full (square) kernel, scale by 1/sum(K), single store. */
static void convolve_to_float(float* out /* w*h*4 */, const oimg* src, int dw, int dh, const float* k, int n)
{
    int r = n / 2;
    float ksum = 0.0f;
    for (int i = 0; i < n * n; ++i) ksum += k[i];
    float scale = 1.0f / ksum;
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < dh; ++y)
        for (int x = 0; x < dw; ++x) {
            float acc[NCH] = { 0, 0, 0, 0 };
            for (int j = 0; j < n; ++j)
                for (int i = 0; i < n; ++i) {
                    float kv = k[j * n + i];
                    float p[NCH];
                    ldpx_clamp(src, x + i - r, y + j - r, p);
                    for (int c = 0; c < NCH; ++c) acc[c] += kv * p[c];
                }
            for (int c = 0; c < NCH; ++c) out[((size_t)y * dw + x) * NCH + c] = scale * acc[c];
        }
}

/* toucan:Blur -- FilterPlugin.cpp:188-202: make_kernel("gaussian", r, r) + convolve */
int o_blur(oimg* dst, const oimg* src, float radius)
{
    int n = o_make_gauss_kernel(radius, NULL, 0);
    float* k = (float*)malloc(sizeof(float) * n * n);
    float* tmp = (float*)malloc(sizeof(float) * (size_t)dst->w * dst->h * NCH);
    if (!k || !tmp) return -1;
    o_make_gauss_kernel(radius, k, n * n);
    convolve_to_float(tmp, src, dst->w, dst->h, k, n);
#pragma omp parallel for schedule(static)
    for (int y = 0; y < dst->h; ++y)
        for (int x = 0; x < dst->w; ++x) stpx(dst, x, y, tmp + ((size_t)y * dst->w + x) * NCH);
    free(k);
    free(tmp);
    return 0;
}

/* toucan:UnsharpMask -- FilterPlugin.cpp:607-618 [V: OIIO unsharp_mask]:
   Blurry(float) = convolve(src, K); Diff = src - Blurry; threshold; Diff *= contrast;
   dst = src + Diff.  Only the "gaussian" family kernel is in scope. */
int o_unsharp(oimg* dst, const oimg* src, float width, float contrast, float threshold)
{
    int n = o_make_gauss_kernel(width, NULL, 0);
    float* k = (float*)malloc(sizeof(float) * n * n);
    float* tmp = (float*)malloc(sizeof(float) * (size_t)dst->w * dst->h * NCH);
    if (!k || !tmp) return -1;
    o_make_gauss_kernel(width, k, n * n);
    convolve_to_float(tmp, src, dst->w, dst->h, k, n);
#pragma omp parallel for schedule(static)
    for (int y = 0; y < dst->h; ++y)
        for (int x = 0; x < dst->w; ++x) {
            float s[NCH], o[NCH];
            ldpx_black(src, x, y, s);
            const float* b = tmp + ((size_t)y * dst->w + x) * NCH;
            for (int c = 0; c < NCH; ++c) {
                float d = s[c] - b[c];
                if (threshold > 0.0f && fabsf(d) < threshold) d = 0.0f;
                d = d * contrast;
                o[c] = s[c] + d;
            }
            stpx(dst, x, y, o);
        }
    free(k);
    free(tmp);
    return 0;
}

/* ------------------------------------------------------------------------- */

/* 9.x draw -- plugins/DrawPlugin.cpp:198-233 (Box), :306-341 (Line): copy(out, src) then
   ImageBufAlgo::render_box / render_line [V: OIIO 3.0 imagebufalgo_draw.cpp].  A drawn pixel is
   "colour over pixel": r[c] = color[c] + r[c] * (1 - alpha), alpha = color[3]; a plain store when
   alpha == 1.  Everything is clipped to the image. */
static void draw_point(oimg* im, int x, int y, const float* color)
{
    if (x < 0 || y < 0 || x >= im->w || y >= im->h) return;
    float p[NCH];
    ldpx(im, x, y, p);
    const float alpha = color[3];
    if (alpha == 1.0f) {
        for (int c = 0; c < NCH; ++c) p[c] = color[c];
    } else {
        const float oma = 1.0f - alpha;
        for (int c = 0; c < NCH; ++c) p[c] = color[c] + p[c] * oma;
    }
    stpx(im, x, y, p);
}
static void copy_into(oimg* dst, const oimg* src)
{
#pragma omp parallel for schedule(static)
    for (int y = 0; y < dst->h; ++y)
        for (int x = 0; x < dst->w; ++x) {
            float p[NCH];
            ldpx_black(src, x, y, p);
            stpx(dst, x, y, p);
        }
}
/* the integer Bresenham walk of OIIO's bresenham2d, first point optionally skipped */
static void draw_line(oimg* im, int x1, int y1, int x2, int y2, const float* color, int skip_first)
{
    int dx = abs(x2 - x1), dy = abs(y2 - y1);
    const int xinc = x1 > x2 ? -1 : 1, yinc = y1 > y2 ? -1 : 1;
    if (dx >= dy) {
        const int dpr = dy << 1, dpru = dpr - (dx << 1);
        int delta = dpr - dx;
        for (; dx >= 0; --dx) {
            if (skip_first) skip_first = 0;
            else draw_point(im, x1, y1, color);
            x1 += xinc;
            if (delta > 0) { y1 += yinc; delta += dpru; }
            else delta += dpr;
        }
    } else {
        const int dpr = dx << 1, dpru = dpr - (dy << 1);
        int delta = dpr - dy;
        for (; dy >= 0; --dy) {
            if (skip_first) skip_first = 0;
            else draw_point(im, x1, y1, color);
            y1 += yinc;
            if (delta > 0) { x1 += xinc; delta += dpru; }
            else delta += dpr;
        }
    }
}
void o_draw_line(oimg* dst, const oimg* src, int x1, int y1, int x2, int y2, const float* color, int skip_first)
{
    copy_into(dst, src);
    draw_line(dst, x1, y1, x2, y2, color, skip_first);
}
void o_draw_box(oimg* dst, const oimg* src, int x1, int y1, int x2, int y2, const float* color, int fill)
{
    copy_into(dst, src);
    if (x1 == x2 && y1 == y2) {
        draw_point(dst, x1, y1, color);
        return;
    }
    if (fill) {
        /* roi_intersection(image, ROI(x1, x2 + 1, y1, y2 + 1)) */
        for (int y = y1 > 0 ? y1 : 0; y < (y2 + 1 < dst->h ? y2 + 1 : dst->h); ++y)
            for (int x = x1 > 0 ? x1 : 0; x < (x2 + 1 < dst->w ? x2 + 1 : dst->w); ++x) draw_point(dst, x, y, color);
        return;
    }
    draw_line(dst, x1, y1, x2, y1, color, 1);
    draw_line(dst, x2, y1, x2, y2, color, 1);
    draw_line(dst, x2, y2, x1, y2, color, 1);
    draw_line(dst, x1, y2, x1, y1, color, 1);
}

/* 9.5 geometry */

/* cut + copy -- TransformPlugin.cpp:197-200: out(x,y) = src(pos+x, pos+y), 0 outside src */
void o_crop(oimg* dst, const oimg* src, int px, int py, int cw, int ch)
{
#pragma omp parallel for schedule(static)
    for (int y = 0; y < dst->h; ++y)
        for (int x = 0; x < dst->w; ++x) {
            float p[NCH] = { 0, 0, 0, 0 };
            if (x < cw && y < ch) ldpx_black(src, px + x, py + y, p);
            stpx(dst, x, y, p);
        }
}
/* flip -- TransformPlugin.cpp:239-246 */
void o_flip(oimg* dst, const oimg* src)
{
#pragma omp parallel for schedule(static)
    for (int y = 0; y < dst->h; ++y)
        for (int x = 0; x < dst->w; ++x) {
            float p[NCH];
            ldpx_black(src, x, src->h - 1 - y, p);
            stpx(dst, x, y, p);
        }
}
/* flop -- TransformPlugin.cpp:284-291 */
void o_flop(oimg* dst, const oimg* src)
{
#pragma omp parallel for schedule(static)
    for (int y = 0; y < dst->h; ++y)
        for (int x = 0; x < dst->w; ++x) {
            float p[NCH];
            ldpx_black(src, src->w - 1 - x, y, p);
            stpx(dst, x, y, p);
        }
}

/* [V: OIIO fmath.h fast_sinpi] */
static inline float fast_sinpi(float x)
{
    const float z = x - ((x + 25165824.0f) - 25165824.0f);
    const float y = z - z * fabsf(z);
    const float Q = 3.10396624f;
    const float P = 3.584135056f;
    return y * (Q + P * fabsf(y));
}
float o_fast_sinpi(float x) { return fast_sinpi(x); }

/* [V: OIIO filter.cpp FilterLanczos3_1D::lanczos3] */
static inline float lanczos3(float x)
{
    const float a = 3.0f;
    const float ainv = 1.0f / a;
    const float m_pi = (float)M_PI;
    x = fabsf(x);
    if (x > a) return 0.0f;
    if (x < 0.0001f) return 1.0f;
    return a / (x * x * (m_pi * m_pi)) * fast_sinpi(x) * fast_sinpi(x * ainv);
}
/* [V: OIIO filter.cpp FilterBlackmanHarris1D::bh1d] */
static inline float bh1d(float x)
{
    if (x < -1.0f || x > 1.0f) return 0.0f;
    x = (x + 1.0f) * 0.5f;
    const float A0 = 0.35875f, A1 = -0.48829f, A2 = 0.14128f, A3 = -0.01168f;
    const float m_pi = (float)M_PI;
    return A0 + A1 * cosf(2.f * m_pi * x) + A2 * cosf(4.f * m_pi * x) + A3 * cosf(6.f * m_pi * x);
}
/* filter id: 0 = lanczos3 (scale = 6/width), 1 = blackman-harris (scale = 2/width) */
static inline float filt1d(int id, float scale, float x) { return id == 0 ? lanczos3(x * scale) : bh1d(x * scale); }

/* Default filter choice of ImageBufAlgo::resize(dst, src, "", 0) [V]:
   blackman-harris (width 3*max(1,ratio)) if either ratio > 1 else lanczos3 (width 6).
   filter_id < 0 selects the default; otherwise filter_width <= 0 uses the filter's
   default width scaled by max(1, ratio). */
void o_resize_filter(int sw, int sh, int dw, int dh, int* id, float* fwx, float* fwy, float user_width)
{
    float wr = (float)dw / (float)sw, hr = (float)dh / (float)sh;
    if (*id < 0) *id = (wr > 1.0f || hr > 1.0f) ? 1 : 0;
    float base = *id == 0 ? 6.0f : 3.0f;
    *fwx = user_width > 0.0f ? user_width : base * (wr > 1.0f ? wr : 1.0f);
    *fwy = user_width > 0.0f ? user_width : base * (hr > 1.0f ? hr : 1.0f);
}

/* Per-axis tap table used by resize_ [V]: for destination index d: source position,
   first tap index and `taps` normalised weights. */
static void resize_axis(int dn, int sn, int id, float fw, int* first, float* wts, int* taps_out, float* tot)
{
    float ratio = (float)dn / (float)sn;
    float rad = fw / 2.0f;
    int radi = (int)ceilf(rad / ratio);
    int taps = 2 * radi + 1;
    *taps_out = taps;
    if (!first) return;
    float scale = id == 0 ? 6.0f / fw : 2.0f / fw;
    float dpix = 1.0f / (float)dn;
    for (int d = 0; d < dn; ++d) {
        float s = ((float)d + 0.5f) * dpix;
        float sf = 0.0f + s * (float)sn;
        float fl = floorf(sf);
        int si = (int)fl;
        float frac = sf - fl;
        float total = 0.0f;
        for (int i = 0; i < taps; ++i) {
            float w = filt1d(id, scale, ratio * ((float)(i - radi) - (frac - 0.5f)));
            wts[(size_t)d * taps + i] = w;
            total += w;
        }
        if (total != 0.0f)
            for (int i = 0; i < taps; ++i) wts[(size_t)d * taps + i] /= total;
        first[d] = si - radi;
        tot[d] = total;
    }
}

/* resize(dst, src, filter, width, ROI(0,w,0,h)) -- TransformPlugin.cpp:374-379,
   Comp.cpp:51-55, TransitionPlugin.cpp:226-230.  [V: OIIO resize_ separable branch]
   filter_id: -1 default, 0 lanczos3, 1 blackman-harris. */
int o_resize(oimg* dst, const oimg* src, int filter_id, float filter_width)
{
    int id = filter_id;
    float fwx, fwy;
    o_resize_filter(src->w, src->h, dst->w, dst->h, &id, &fwx, &fwy, filter_width);
    int xt, yt;
    resize_axis(dst->w, src->w, id, fwx, NULL, NULL, &xt, NULL);
    resize_axis(dst->h, src->h, id, fwy, NULL, NULL, &yt, NULL);
    int* xf = (int*)malloc(sizeof(int) * dst->w);
    int* yf = (int*)malloc(sizeof(int) * dst->h);
    float* xw = (float*)malloc(sizeof(float) * (size_t)dst->w * xt);
    float* yw = (float*)malloc(sizeof(float) * (size_t)dst->h * yt);
    float* xtot = (float*)malloc(sizeof(float) * dst->w);
    float* ytot = (float*)malloc(sizeof(float) * dst->h);
    if (!xf || !yf || !xw || !yw || !xtot || !ytot) return -1;
    resize_axis(dst->w, src->w, id, fwx, xf, xw, &xt, xtot);
    resize_axis(dst->h, src->h, id, fwy, yf, yw, &yt, ytot);
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < dst->h; ++y)
        for (int x = 0; x < dst->w; ++x) {
            float pel[NCH] = { 0, 0, 0, 0 };
            /* totalweight_x is re-summed from the already normalised taps in OIIO */
            float tx = 0.0f;
            for (int i = 0; i < xt; ++i) tx += xw[(size_t)x * xt + i];
            if (tx != 0.0f) {
                for (int j = 0; j < yt; ++j) {
                    float wy = yw[(size_t)y * yt + j];
                    if (wy == 0.0f) continue;
                    for (int i = 0; i < xt; ++i) {
                        float w = wy * xw[(size_t)x * xt + i];
                        if (w != 0.0f) {
                            float p[NCH];
                            ldpx_clamp(src, xf[x] + i, yf[y] + j, p);
                            for (int c = 0; c < NCH; ++c) pel[c] += w * p[c];
                        }
                    }
                }
            }
            if (ytot[y] == 0.0f) pel[0] = pel[1] = pel[2] = pel[3] = 0.0f;
            stpx(dst, x, y, pel);
        }
    free(xf); free(yf); free(xw); free(yw); free(xtot); free(ytot);
    return 0;
}

/* Inverse rotation matrix used by rotate() [V: OIIO rotate -> warp; Imath M33f]:
   M = T(-c) * R(angle) * T(c) in Imath's row-vector convention with
   R = [[cos, sin],[-sin, cos]], centre = centre of the source full window.
   Minv (row-major 3x3, row-vector convention) is written to m[9]. */
void o_rotate_matrix(int w, int h, float angle, float* minv)
{
    float cx = 0.5f * (float)(0 + w), cy = 0.5f * (float)(0 + h);
    float c = cosf(angle), s = sinf(angle);
    /* M = T(-c) R T(c):  rows */
    float M[3][3] = { { c, s, 0.0f }, { -s, c, 0.0f }, { 0.0f, 0.0f, 1.0f } };
    /* T(-c) * R : third row = (-cx, -cy, 1) * R */
    M[2][0] = -cx * c + -cy * -s;
    M[2][1] = -cx * s + -cy * c;
    /* * T(c) */
    M[2][0] += cx;
    M[2][1] += cy;
    /* inverse of an affine row-vector matrix the way Imath's Matrix33::inverse() does
       it for the affine case (x[0][2]==x[1][2]==0, x[2][2]==1): 2x2 adjugate / det,
       then the translation row from the already divided entries. */
    float r = M[0][0] * M[1][1] - M[1][0] * M[0][1];
    float inv[3][3];
    inv[0][0] = M[1][1] / r;
    inv[0][1] = -M[0][1] / r;
    inv[0][2] = 0.0f;
    inv[1][0] = -M[1][0] / r;
    inv[1][1] = M[0][0] / r;
    inv[1][2] = 0.0f;
    inv[2][0] = -M[2][0] * inv[0][0] - M[2][1] * inv[1][0];
    inv[2][1] = -M[2][0] * inv[0][1] - M[2][1] * inv[1][1];
    inv[2][2] = 1.0f;
    memcpy(minv, inv, sizeof(inv));
}

/* rotate(dst, src, angle, "", 0) -- TransformPlugin.cpp:462-467.
   [V: OIIO warp_ + filtered_sample]: inverse map of the pixel centre, lanczos3
   (width 6, 2-D separable product) footprint floor(s-3)..ceil(s+3) exclusive, black
   outside the source, result = sum(w p) / sum(w) (weights of outside taps count). */
void o_rotate(oimg* dst, const oimg* src, float angle)
{
    float mi[9];
    o_rotate_matrix(src->w, src->h, angle, mi);
    const float fw = 6.0f;
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < dst->h; ++y)
        for (int x = 0; x < dst->w; ++x) {
            float px = (float)x + 0.5f, py = (float)y + 0.5f;
            float s = px * mi[0] + py * mi[3] + mi[6];
            float t = px * mi[1] + py * mi[4] + mi[7];
            float dsdx = mi[0], dsdy = mi[3], dtdx = mi[1], dtdy = mi[4];
            float ds = fmaxf(1.0f, fmaxf(fabsf(dsdx), fabsf(dsdy)));
            float dt = fmaxf(1.0f, fmaxf(fabsf(dtdx), fabsf(dtdy)));
            float ds_inv = 1.0f / ds, dt_inv = 1.0f / dt;
            float rs = 0.5f * ds * fw, rt = 0.5f * dt * fw;
            int smin = (int)floorf(s - rs), smax = (int)ceilf(s + rs);
            int tmin = (int)floorf(t - rt), tmax = (int)ceilf(t + rt);
            float sum[NCH] = { 0, 0, 0, 0 };
            float total = 0.0f;
            for (int j = tmin; j < tmax; ++j)
                for (int i = smin; i < smax; ++i) {
                    float w = lanczos3(ds_inv * ((float)i + 0.5f - s) * (6.0f / fw)) *
                              lanczos3(dt_inv * ((float)j + 0.5f - t) * (6.0f / fw));
                    float p[NCH];
                    ldpx_black(src, i, j, p);
                    for (int c = 0; c < NCH; ++c) sum[c] += w * p[c];
                    total += w;
                }
            float o[NCH] = { 0, 0, 0, 0 };
            if (total > 0.0f)
                for (int c = 0; c < NCH; ++c) o[c] = sum[c] / total;
            stpx(dst, x, y, o);
        }
}

/* ------------------------------------------------------------------------- */
/* 9.6 compositing and transitions (toucan code, exact) */

/* paste(dst, ox, oy, src): copy src into dst at offset, clipped, with type conversion */
void o_paste(oimg* dst, int ox, int oy, const oimg* src)
{
#pragma omp parallel for schedule(static)
    for (int y = 0; y < src->h; ++y) {
        int dy = oy + y;
        if (dy < 0 || dy >= dst->h) continue;
        for (int x = 0; x < src->w; ++x) {
            int dx = ox + x;
            if (dx < 0 || dx >= dst->w) continue;
            float p[NCH];
            ldpx(src, x, y, p);
            stpx(dst, dx, dy, p);
        }
    }
}

/* Util.cpp:123-147 fit(): box of b fitted into a, integer truncation.  out = {minx,miny,w,h} */
void o_fit(int ax, int ay, int bx, int by, int* out)
{
    int w = bx, h = by;
    double aAspect = ay > 0 ? (ax / (double)ay) : 1.0;
    double bAspect = by > 0 ? (bx / (double)by) : 1.0;
    if (bAspect > aAspect) {
        w = ax;
        h = (int)(w / bAspect);
    } else {
        h = ay;
        w = (int)(h * bAspect);
    }
    out[0] = ax / 2 - w / 2;
    out[1] = ay / 2 - h / 2;
    out[2] = w;
    out[3] = h;
}

/* over(fg, bg) [V: OIIO over_impl]: alpha = clamp(fg.a,0,1); out = fg + (1-alpha)*bg,
   all four channels.  fg must already be premultiplied and same size as bg. */
void o_over(oimg* dst, const oimg* fg, const oimg* bg)
{
#pragma omp parallel for schedule(static)
    for (int y = 0; y < dst->h; ++y)
        for (int x = 0; x < dst->w; ++x) {
            float a[NCH], b[NCH], o[NCH];
            ldpx_black(fg, x, y, a);
            ldpx_black(bg, x, y, b);
            float alpha = o_clampf(a[3], 0.0f, 1.0f);
            float oma = 1.0f - alpha;
            for (int c = 0; c < NCH; ++c) o[c] = a[c] + oma * b[c];
            stpx(dst, x, y, o);
        }
}

/* Dissolve -- TransitionPlugin.cpp:242-264, sizes already equal:
   A = mul(from, fill(iv))  (typed like `from`: the constant and the product are each
   quantised to from's storage type), B = mul(to, fill(v)) (typed like `to`),
   out = add(A, B) stored in dst's type. */
void o_dissolve(oimg* dst, const oimg* from, const oimg* to, double value)
{
    const float v = (float)value;
    const float iv = (float)(1.0 - value);
    const float qiv = quant(iv, from->type), qv = quant(v, to->type);
#pragma omp parallel for schedule(static)
    for (int y = 0; y < dst->h; ++y)
        for (int x = 0; x < dst->w; ++x) {
            float a[NCH], b[NCH], o[NCH];
            ldpx_black(from, x, y, a);
            ldpx_black(to, x, y, b);
            for (int c = 0; c < NCH; ++c) {
                float A = quant(a[c] * qiv, from->type);
                float B = quant(b[c] * qv, to->type);
                o[c] = A + B;
            }
            stpx(dst, x, y, o);
        }
}

/* wipe matte value at offset k = px - edge along the wipe axis (u8 RGBA mattes,
   TransitionPlugin.cpp:309-338 / 399-428):  Mf: 0 | q(k/199) on [0,200) | 1;
   Mt: 1 | q((199-k)/199) on [0,200) | 0. */
static inline void wipe_mattes(int k, float* mf, float* mt)
{
    if (k < 0) { *mf = 0.0f; *mt = 1.0f; }
    else if (k < 200) {
        /* fill(gradient, 0, 1) on a 200-row u8 image: row r -> lerp(0,1,r/199) */
        *mf = quant(o_lerp(0.0f, 1.0f, (float)k / 199.0f), O_U8);
        *mt = quant(o_lerp(0.0f, 1.0f, (float)(199 - k) / 199.0f), O_U8);
    } else { *mf = 1.0f; *mt = 0.0f; }
}

/* HorizontalWipe (vertical=0) / VerticalWipe (vertical=1):
   out = add(mul(from, Mf), mul(to, Mt)); mul results are typed merge(src, u8) = src type. */
void o_wipe(oimg* dst, const oimg* from, const oimg* to, double value, int vertical)
{
    const int w = from->w, h = from->h;
    const int edge = vertical ? (int)(h * value) : (int)(w * value);
#pragma omp parallel for schedule(static)
    for (int y = 0; y < dst->h; ++y)
        for (int x = 0; x < dst->w; ++x) {
            float a[NCH], b[NCH], o[NCH], mf, mt;
            /* mattes are w x h (from's size): outside them the matte reads 0 */
            if (x < w && y < h) wipe_mattes((vertical ? y : x) - edge, &mf, &mt);
            else mf = mt = 0.0f;
            ldpx_black(from, x, y, a);
            ldpx_black(to, x, y, b);
            for (int c = 0; c < NCH; ++c) {
                float A = quant(a[c] * mf, from->type);
                float B = quant(b[c] * mt, to->type);
                o[c] = A + B;
            }
            stpx(dst, x, y, o);
        }
}

/* ------------------------------------------------------------------------- */
/* 9.7 ColorConvert: baked curve -> 3x3 -> curve.  [V: OCIO 2.4.2 built-in CG config]
   curve ids: 0 linear, 1 gamma (param = exponent, pass-thru negatives), 2 sRGB
   piecewise, 3 ACEScct, 4 ACEScc; dir 0 = decode (to linear), 1 = encode. */
static inline float cc_curve(int id, int encode, float p, float x)
{
    switch (id) {
    case 1:
        if (x <= 0.0f) return x;
        return powf(x, encode ? 1.0f / p : p);
    case 2:
        if (!encode) return x <= 0.04045f ? x / 12.92f : powf((x + 0.055f) / 1.055f, 2.4f);
        return x <= 0.0031308f ? x * 12.92f : 1.055f * powf(x, 1.0f / 2.4f) - 0.055f;
    case 3:
        if (encode) return x <= 0.0078125f ? 10.5402377416545f * x + 0.0729055341958355f : (log2f(x) + 9.72f) / 17.52f;
        return x <= 0.155251141552511f ? (x - 0.0729055341958355f) / 10.5402377416545f : exp2f(x * 17.52f - 9.72f);
    case 4:
        if (encode) {
            if (x <= 0.0f) return -0.3584474886f;
            if (x < 0.000030517578125f) return (log2f(0.0000152587890625f + x * 0.5f) + 9.72f) / 17.52f;
            return (log2f(x) + 9.72f) / 17.52f;
        }
        if (x <= -0.3013698630f) return (exp2f(x * 17.52f - 9.72f) - 0.0000152587890625f) * 2.0f;
        return exp2f(x * 17.52f - 9.72f);
    default: return x;
    }
}
/* unpremult != 0 [V: OIIO colorconvert(..., unpremult, ...)]: pixels with alpha != 0 are divided by alpha
   before the transform and multiplied by it afterwards (toucan passes its "premult" parameter here,
   ColorPlugin.cpp:241; its effective default is 0). */
void o_colorconvert2(oimg* dst, const oimg* src, int curve_in, float p_in, const float* m33, int curve_out, float p_out, int unpremult)
{
#pragma omp parallel for schedule(static)
    for (int y = 0; y < dst->h; ++y)
        for (int x = 0; x < dst->w; ++x) {
            float p[NCH], l[3], o[NCH];
            ldpx_black(src, x, y, p);
            const float a = p[3];
            const int un = unpremult && a != 0.0f && a != 1.0f;
            if (un)
                for (int c = 0; c < 3; ++c) p[c] = p[c] / a;
            for (int c = 0; c < 3; ++c) l[c] = cc_curve(curve_in, 0, p_in, p[c]);
            for (int c = 0; c < 3; ++c) {
                float v = (m33[c * 3 + 0] * l[0] + m33[c * 3 + 1] * l[1]) + m33[c * 3 + 2] * l[2];
                o[c] = cc_curve(curve_out, 1, p_out, v);
            }
            if (un)
                for (int c = 0; c < 3; ++c) o[c] = o[c] * a;
            o[3] = a;
            stpx(dst, x, y, o);
        }
}
void o_colorconvert(oimg* dst, const oimg* src, int curve_in, float p_in, const float* m33, int curve_out, float p_out)
{
#pragma omp parallel for schedule(static)
    for (int y = 0; y < dst->h; ++y)
        for (int x = 0; x < dst->w; ++x) {
            float p[NCH], l[3], o[NCH];
            ldpx_black(src, x, y, p);
            for (int c = 0; c < 3; ++c) l[c] = cc_curve(curve_in, 0, p_in, p[c]);
            for (int c = 0; c < 3; ++c) {
                float v = (m33[c * 3 + 0] * l[0] + m33[c * 3 + 1] * l[1]) + m33[c * 3 + 2] * l[2];
                o[c] = cc_curve(curve_out, 1, p_out, v);
            }
            o[3] = p[3];
            stpx(dst, x, y, o);
        }
}

/* ------------------------------------------------------------------------- */
/* storage conversion / copy with black outside (paste(tmp,0,0,buf) of App.cpp:285-291) */
void o_convert(oimg* dst, const oimg* src)
{
#pragma omp parallel for schedule(static)
    for (int y = 0; y < dst->h; ++y)
        for (int x = 0; x < dst->w; ++x) {
            float p[NCH];
            ldpx_black(src, x, y, p);
            stpx(dst, x, y, p);
        }
}
