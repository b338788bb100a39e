// tc_pointwise.cu -- every per-pixel effect of the toucan render path as ONE pass over
// HBM: generators, pointwise filters, address-only transforms, transitions and the
// premult+over compositor.  Compiled with -fmad=false: the reference (OIIO on baseline
// x86-64) has no fused multiply-add, and Fill/Checkers/Invert/Flip/Flop/Crop/Wipe/Noise
// must be bit-exact.
//
// Kernel shape: grid-stride over pixels, one RGBA pixel per thread-iteration (128-bit
// load/store for float, 64-bit for half/u16, 32-bit for u8), 4 independent pixels in
// flight per thread, grid sized as a multiple of the SM count.
#include <algorithm>
#include <cstring>
#include <mutex>
#include <vector>

#include "tc_common.cuh"

namespace tc {

constexpr int PW_THREADS = 256;
constexpr int PW_PIXELS = 4; // pixels per thread-iteration

// VEC pixels per 128-bit access: 1 (float), 2 (half / u16), 4 (u8).  Every thread-iteration moves
// PW_PIXELS / VEC independent 16-byte groups, so small storage types get the same bytes in flight as float.
template <int VEC>
__device__ __forceinline__ void store_vec(void* __restrict__ base, int type, unsigned idx0, const float4 (&v)[VEC])
{
    if (VEC == 1) {
        store_px(base, type, idx0, v[0]);
    } else if (VEC == 4) {
        uint4 o;
        unsigned* w = &o.x;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            w[j] = f_to_u8(v[j].x) | (f_to_u8(v[j].y) << 8) | (f_to_u8(v[j].z) << 16) | (f_to_u8(v[j].w) << 24);
        reinterpret_cast<uint4*>(base)[idx0 >> 2] = o;
    } else {
        uint4 o;
        unsigned* w = &o.x;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            if (type == TC_U16) {
                w[2 * j] = f_to_u16(v[j].x) | (f_to_u16(v[j].y) << 16);
                w[2 * j + 1] = f_to_u16(v[j].z) | (f_to_u16(v[j].w) << 16);
            } else {
                __half2 a = __floats2half2_rn(v[j].x, v[j].y), b = __floats2half2_rn(v[j].z, v[j].w);
                w[2 * j] = *reinterpret_cast<unsigned*>(&a);
                w[2 * j + 1] = *reinterpret_cast<unsigned*>(&b);
            }
        }
        reinterpret_cast<uint4*>(base)[idx0 >> 1] = o;
    }
}

template <class Op, int VEC>
__global__ void __launch_bounds__(PW_THREADS) k_pointwise(const __grid_constant__ Op op, const Img dst)
{
    constexpr int PW_UNROLL = PW_PIXELS / VEC; // independent 128-bit groups per thread-iteration
    const unsigned groups = ((unsigned)dst.w * (unsigned)dst.h) / VEC;
    const unsigned stride = gridDim.x * blockDim.x;
    for (unsigned base = blockIdx.x * blockDim.x + threadIdx.x; base < groups; base += PW_UNROLL * stride) {
        float4 v[PW_UNROLL][VEC];
        typename Op::template Cache<VEC> cache[PW_UNROLL];
#pragma unroll
        for (int k = 0; k < PW_UNROLL; ++k) {
            const unsigned g = base + k * stride;
            if (g < groups) op.template preload<VEC>(cache[k], g * VEC);
        }
#pragma unroll
        for (int k = 0; k < PW_UNROLL; ++k) {
            const unsigned g = base + k * stride;
            if (g < groups) {
                const unsigned idx0 = g * VEC;
                int y = idx0 / (unsigned)dst.w;
                int x = idx0 - y * (unsigned)dst.w;
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    v[k][j] = op.template eval<VEC>(cache[k], x, y, idx0 + j, j);
                    if (++x == dst.w) {
                        x = 0;
                        ++y;
                    }
                }
            }
        }
#pragma unroll
        for (int k = 0; k < PW_UNROLL; ++k) {
            const unsigned g = base + k * stride;
            if (g < groups) store_vec<VEC>(dst.p, dst.type, g * VEC, v[k]);
        }
    }
}

template <class Op, int VEC>
static int launch_vec(const Op& op, const tc_image* dst, tc_stream s)
{
    const size_t groups = (size_t)dst->w * dst->h / VEC;
    const size_t per_block = (size_t)PW_THREADS * (PW_PIXELS / VEC);
    size_t blocks = (groups + per_block - 1) / per_block;
    const size_t cap = (size_t)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    k_pointwise<Op, VEC><<<(unsigned)blocks, PW_THREADS, 0, (cudaStream_t)s>>>(op, view(dst));
    count_launch();
    return check_cuda(cudaGetLastError(), "k_pointwise launch");
}

template <class Op>
static int launch(const Op& op, const tc_image* dst, tc_stream s)
{
    const size_t total = (size_t)dst->w * dst->h;
    if (total >= (1ull << 31)) return fail(TC_ERR_UNSUPPORTED, "image too large (%d x %d)", dst->w, dst->h);
    // 128-bit groups of the destination's pixels; sources of any storage type (Src::preload)
    const int vec = dst->type == TC_F32 ? 1 : (dst->type == TC_U8 ? 4 : 2);
    if (vec > 1 && total % vec == 0 && op.vec_ok(dst->type)) {
        if (vec == 4) return launch_vec<Op, 4>(op, dst, s);
        return launch_vec<Op, 2>(op, dst, s);
    }
    return launch_vec<Op, 1>(op, dst, s);
}

// A source image read at the destination's pixel position (0 outside).  In the vectorised kernels a thread owns a
// group of VEC consecutive pixels (VEC comes from the DESTINATION's storage type); a same-sized source supplies those
// pixels as one 32/64/128-bit word when VEC of ITS pixels fit in 16 bytes -- whatever its own type, so a half frame
// over a u8 frame is vectorised too -- and as per-pixel loads (each at least 64 bits wide) otherwise.
__host__ __device__ __forceinline__ int px_bytes(int type) { return type == TC_U8 ? 4 : (type == TC_F32 ? 16 : 8); }

struct Src {
    Img im;
    int same; // same dimensions as the destination -> linear index can be reused
    bool vec_ok(int) const { return true; }
    template <int VEC>
    __device__ __forceinline__ void preload(uint4& c, unsigned idx0) const
    {
        if (VEC > 1 && same) {
            const int bytes = VEC * px_bytes(im.type);
            const char* p = reinterpret_cast<const char*>(im.p) + (size_t)idx0 * px_bytes(im.type);
            if (bytes == 16) {
                c = __ldg(reinterpret_cast<const uint4*>(p));
            } else if (bytes == 8) {
                const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
                c.x = v.x;
                c.y = v.y;
            } else if (bytes == 4) {
                c.x = __ldg(reinterpret_cast<const unsigned*>(p));
            }
        }
    }
    template <int VEC>
    __device__ __forceinline__ float4 at(const uint4& c, int x, int y, unsigned idx, int j) const
    {
        if (VEC == 1) return same ? load_px(im.p, im.type, idx) : load_black(im, x, y);
        if (!same) return load_black(im, x, y);
        if (VEC * px_bytes(im.type) > 16) return load_px(im.p, im.type, idx);
        const unsigned* w = &c.x;
        if (im.type == TC_U8) {
            const unsigned v = w[j];
            return make_float4(u8_to_f(v & 0xffu), u8_to_f((v >> 8) & 0xffu), u8_to_f((v >> 16) & 0xffu), u8_to_f(v >> 24));
        }
        const unsigned lo = w[(2 * j) & 3], hi = w[(2 * j + 1) & 3]; // two 16-bit pixels per word pair (VEC <= 2 here)
        if (im.type == TC_U16) return make_float4(u16_to_f(lo & 0xffffu), u16_to_f(lo >> 16), u16_to_f(hi & 0xffffu), u16_to_f(hi >> 16));
        const float2 fa = __half22float2(*reinterpret_cast<const __half2*>(&lo)), fb = __half22float2(*reinterpret_cast<const __half2*>(&hi));
        return make_float4(fa.x, fa.y, fb.x, fb.y);
    }
};
static Src make_src(const tc_image* src, const tc_image* dst)
{
    return Src { view(src), (src->w == dst->w && src->h == dst->h) ? 1 : 0 };
}

// Boilerplate of an op: N same-position sources (0, 1 or 2) cached as one 128-bit word each.
struct NoCache {};
#define TC_OP_0()                                                                             \
    template <int VEC> using Cache = NoCache;                                                 \
    bool vec_ok(int) const { return true; }                                                   \
    template <int VEC> __device__ __forceinline__ void preload(NoCache&, unsigned) const {}   \
    template <int VEC> __device__ __forceinline__ float4 eval(const NoCache&, int x, int y, unsigned i, int) const { return (*this)(x, y, i); }
struct Cache1 { uint4 a; };
struct Cache2 { uint4 a, b; };
#define TC_OP_1(S)                                                                                      \
    template <int VEC> using Cache = Cache1;                                                            \
    bool vec_ok(int t) const { return S.vec_ok(t); }                                                    \
    template <int VEC> __device__ __forceinline__ void preload(Cache1& c, unsigned i0) const { S.template preload<VEC>(c.a, i0); } \
    template <int VEC> __device__ __forceinline__ float4 eval(const Cache1& c, int x, int y, unsigned i, int j) const { return apply(S.template at<VEC>(c.a, x, y, i, j)); }
#define TC_OP_2(SA, SB)                                                                                 \
    template <int VEC> using Cache = Cache2;                                                            \
    bool vec_ok(int t) const { return SA.vec_ok(t) && SB.vec_ok(t); }                                   \
    template <int VEC> __device__ __forceinline__ void preload(Cache2& c, unsigned i0) const { SA.template preload<VEC>(c.a, i0); SB.template preload<VEC>(c.b, i0); } \
    template <int VEC> __device__ __forceinline__ float4 eval(const Cache2& c, int x, int y, unsigned i, int j) const { return apply(SA.template at<VEC>(c.a, x, y, i, j), SB.template at<VEC>(c.b, x, y, i, j), x, y); }

__device__ __forceinline__ float lerp_oiio(float a, float b, float x)
{
    // OIIO fmath.h lerp: v0*(1-x) + v1*x
    return __fadd_rn(__fmul_rn(a, __fsub_rn(1.0f, x)), __fmul_rn(b, x));
}
__device__ __forceinline__ float luma709(float4 p)
{
    return __fadd_rn(__fadd_rn(__fmul_rn(0.2126f, p.x), __fmul_rn(0.7152f, p.y)), __fmul_rn(0.0722f, p.z));
}

// ---- generators ------------------------------------------------------------------------
struct FillOp {
    TC_OP_0()
    float4 c;
    __device__ float4 operator()(int, int, unsigned) const { return c; }
};
struct CheckerOp {
    TC_OP_0()
    float4 c1, c2;
    int cw, ch;
    unsigned mx, my; // floor(2^32 / c) + 1: x / c == umulhi(x, m) exactly while x * c < 2^32 (0 = divide)
    __device__ float4 operator()(int x, int y, unsigned) const
    {
        const unsigned tx = mx ? __umulhi((unsigned)x, mx) : (unsigned)(x / cw);
        const unsigned ty = my ? __umulhi((unsigned)y, my) : (unsigned)(y / ch);
        return ((tx + ty) & 1) ? c2 : c1;
    }
};
struct GradientOp {
    TC_OP_0()
    float4 c1, c2;
    float denom;
    int vertical;
    __device__ float4 operator()(int x, int y, unsigned) const
    {
        const float v = __fdiv_rn((float)(vertical ? y : x), denom);
        float4 r = make_float4(lerp_oiio(c1.x, c2.x, v), lerp_oiio(c1.y, c2.y, v), lerp_oiio(c1.z, c2.z, v),
                               lerp_oiio(c1.w, c2.w, v));
        // horizontal: the reference ramps a UINT8 temporary, rotates and copies it
        return vertical ? r : quant4(r, TC_U8);
    }
};

__device__ __forceinline__ uint32_t rotl32(uint32_t x, int k) { return (x << k) | (x >> (32 - k)); }
__device__ __forceinline__ uint32_t bjfinal(uint32_t a, uint32_t b, uint32_t c)
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
__device__ __forceinline__ float hashrand(uint32_t xy, int c, int seed)
{
    const uint32_t h = bjfinal(xy, (uint32_t)c, (uint32_t)seed) & 0xfffffu;
    return __fmul_rn((float)h, 1.0f / 1048576.0f);
}
// ln(x) in double from + - * / only (the fdlibm e_log scheme), so host restatements of the
// same sequence agree to the bit.  Callers pass normal positive finite x.
__device__ double det_log(double x)
{
    const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10,
                 Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01,
                 Lg4 = 2.222219843214978396e-01, Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
                 Lg7 = 1.479819860511658591e-01;
    int hx = __double2hiint(x);
    const int lx = __double2loint(x);
    int k = (hx >> 20) - 1023;
    hx &= 0x000fffff;
    int i = (hx + 0x95f64) & 0x100000;
    x = __hiloint2double(hx | (i ^ 0x3ff00000), lx);
    k += (i >> 20);
    const double f = __dsub_rn(x, 1.0), dk = (double)k;
    if ((0x000fffff & (2 + hx)) < 3) {
        if (f == 0.0) {
            if (k == 0) return 0.0;
            return __dadd_rn(__dmul_rn(dk, ln2_hi), __dmul_rn(dk, ln2_lo));
        }
        const double R = __dmul_rn(__dmul_rn(f, f), __dsub_rn(0.5, __dmul_rn(0.33333333333333333, f)));
        if (k == 0) return __dsub_rn(f, R);
        return __dsub_rn(__dmul_rn(dk, ln2_hi), __dsub_rn(__dsub_rn(R, __dmul_rn(dk, ln2_lo)), f));
    }
    const double s = __ddiv_rn(f, __dadd_rn(2.0, f));
    const double z = __dmul_rn(s, s);
    i = hx - 0x6147a;
    const double w = __dmul_rn(z, z);
    const int j = 0x6b851 - hx;
    const double t1 = __dmul_rn(w, __dadd_rn(Lg2, __dmul_rn(w, __dadd_rn(Lg4, __dmul_rn(w, Lg6)))));
    const double t2 = __dmul_rn(
        z, __dadd_rn(Lg1, __dmul_rn(w, __dadd_rn(Lg3, __dmul_rn(w, __dadd_rn(Lg5, __dmul_rn(w, Lg7)))))));
    i |= j;
    const double R = __dadd_rn(t2, t1);
    if (i > 0) {
        const double hfsq = __dmul_rn(__dmul_rn(0.5, f), f);
        if (k == 0) return __dsub_rn(f, __dsub_rn(hfsq, __dmul_rn(s, __dadd_rn(hfsq, R))));
        return __dsub_rn(__dmul_rn(dk, ln2_hi),
                         __dsub_rn(__dsub_rn(hfsq, __dadd_rn(__dmul_rn(s, __dadd_rn(hfsq, R)), __dmul_rn(dk, ln2_lo))), f));
    }
    if (k == 0) return __dsub_rn(f, __dmul_rn(s, __dsub_rn(f, R)));
    return __dsub_rn(__dmul_rn(dk, ln2_hi), __dsub_rn(__dsub_rn(__dmul_rn(s, __dsub_rn(f, R)), __dmul_rn(dk, ln2_lo)), f));
}
// sqrt(-2 ln(x) / x) for x in (0, 1], FP64 with fused multiply-adds and Newton-refined SFU seeds (see hashnormal): the same
// fdlibm decomposition as det_log, the quotient f / (2 + f) through a refined reciprocal.
__device__ __forceinline__ double fast_rcp_d(double b)
{
    double r = (double)rcp_approx((float)b);
    r = fma(r, fma(-b, r, 1.0), r);
    r = fma(r, fma(-b, r, 1.0), r);
    return r;
}
__device__ __forceinline__ double fast_noise_scale(double x0)
{
    const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10,
                 Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01,
                 Lg4 = 2.222219843214978396e-01, Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
                 Lg7 = 1.479819860511658591e-01;
    int hx = __double2hiint(x0);
    const int lx = __double2loint(x0);
    int k = (hx >> 20) - 1023;
    hx &= 0x000fffff;
    const int i = (hx + 0x95f64) & 0x100000;
    const double x = __hiloint2double(hx | (i ^ 0x3ff00000), lx);
    k += (i >> 20);
    const double f = x - 1.0, dk = (double)k;
    const double s = f * fast_rcp_d(2.0 + f);
    const double z = s * s, w = z * z;
    const double t1 = w * fma(w, fma(w, Lg6, Lg4), Lg2);
    const double t2 = z * fma(w, fma(w, fma(w, Lg7, Lg5), Lg3), Lg1);
    const double R = t2 + t1;
    const double hfsq = 0.5 * f * f;
    // ln(x0) = k ln2 + f - (hfsq - s (hfsq + R)), the low parts first
    const double L = fma(dk, ln2_hi, -((hfsq - fma(s, hfsq + R, dk * ln2_lo)) - f));
    const double q = (-2.0 * L) * fast_rcp_d(x0);
    double y = (double)rsqrt_approx((float)q);
    y = y * fma(-0.5 * q * y, y, 1.5);
    y = y * fma(-0.5 * q * y, y, 1.5);
    const double m = q * y;
    return fma(fma(-m, m, q), 0.5 * y, m);
}
// Marsaglia polar method on the position hash: candidates (s, s + 139), s = seed, seed + 1, ... until one lies inside the unit
// circle.  marsaglia_scale finishes an accepted candidate.
__device__ __forceinline__ float marsaglia_scale(float xr, float r2)
{
    // M = (float)sqrt(-2 log(r2) / r2), every step an IEEE double operation in the reference.  Only the FLOAT M is used, so
    // the double chain is evaluated with reciprocal / rsqrt seeds + Newton steps and fused multiply-adds (~50 FP64
    // instructions instead of ~180: IEEE division twice, sqrt and the unfused log); its result is within 2 double ulps of
    // the exact chain (tests/cpp/noise_fastpath_check.c sweeps every float r2), and whenever it lands within 1024 ulps of a float
    // rounding boundary -- one sample in 2^18 -- or is not a positive normal, the exact chain decides.  Bit-identical to it.
    const double r2d = (double)r2;
    const double Mf = fast_noise_scale(r2d);
    const unsigned low = (unsigned)__double2loint(Mf) & 0x1fffffffu; // the 29 bits below float precision
    const unsigned dist = low > 0x10000000u ? low - 0x10000000u : 0x10000000u - low;
    float M;
    if (dist < 1024u || !(Mf > 1.0e-30)) {
        const double q = __ddiv_rn(__dmul_rn(-2.0, det_log(r2d)), r2d);
        M = (float)__dsqrt_rn(q);
    } else
        M = (float)Mf;
    return __fmul_rn(xr, M);
}
// The accepted candidates of a pixel's nc channels.  One flat loop over (channel, try) instead of a rejection loop per
// channel: a warp runs as long as its slowest lane, and the sum of four geometric counts spreads far less than four maxima
// (8 instead of 13.6 hash rounds per pixel on average).  The FP64 finish runs after the loop, undiverged, four chains deep.
__device__ __forceinline__ void hashnormal4(uint32_t xy, int seed, int nc, float (&out)[4])
{
    float xa[4] = { 0.f, 0.f, 0.f, 0.f }, ra[4] = { 1.f, 1.f, 1.f, 1.f };
    int c = 0, s = seed;
    while (c < nc) {
        // (float)(2.0 * (double)h - 1.0): h is k / 2^20, so 2h - 1 has at most 21 significant bits and the float
        // evaluation is exact too (checked for all 2^20 values) -- no trip through the conversion and FP64 pipes
        const float xr = __fsub_rn(__fmul_rn(2.0f, hashrand(xy, c, s)), 1.0f);
        const float yr = __fsub_rn(__fmul_rn(2.0f, hashrand(xy, c, s + 139)), 1.0f);
        const float r2 = __fadd_rn(__fmul_rn(xr, xr), __fmul_rn(yr, yr));
        if (r2 > 1.0f || r2 == 0.0f) {
            ++s;
            continue;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (c == k) xa[k] = xr, ra[k] = r2;
        ++c;
        s = seed;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) out[k] = k < nc ? marsaglia_scale(xa[k], ra[k]) : 0.0f;
}
struct NoiseOp {
    TC_OP_0()
    float A, B;
    int type, mono, seed;
    __device__ float4 operator()(int x, int y, unsigned) const
    {
        const uint32_t xy = bjfinal((uint32_t)x, (uint32_t)y, 0u);
        float p[4] = { 0.f, 0.f, 0.f, 0.f }; // the output buffer is zero-initialised (ImageEffect.cpp:93)
        float n = 0.f;
        float g[4] = { 0.f, 0.f, 0.f, 0.f };
        if (type == 0) hashnormal4(xy, seed, mono ? 1 : 4, g);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (type == 0) {
                if (c == 0 || !mono) n = __fadd_rn(A, __fmul_rn(B, g[c]));
                p[c] = __fadd_rn(p[c], n);
            } else if (type == 1) {
                if (c == 0 || !mono) n = __fadd_rn(A, __fmul_rn(__fsub_rn(B, A), hashrand(xy, c, seed)));
                p[c] = __fadd_rn(p[c], n);
            } else {
                if (c == 0 || !mono) n = hashrand(xy, c, seed);
                if (n < B) p[c] = A;
            }
        }
        return make_float4(p[0], p[1], p[2], p[3]);
    }
};

// ---- pointwise filters -----------------------------------------------------------------
struct InvertOp {
    Src s;
    TC_OP_1(s)
    __device__ __forceinline__ float4 apply(float4 p) const
    {
        p.x = __fadd_rn(__fmul_rn(p.x, -1.0f), 1.0f);
        p.y = __fadd_rn(__fmul_rn(p.y, -1.0f), 1.0f);
        p.z = __fadd_rn(__fmul_rn(p.z, -1.0f), 1.0f);
        return p;
    }
};
// powf for the tolerance path.  CUDA's powf is ~95 instructions (double-float log2 and a tree of special
// cases); this one is ~40 and good to ~4 ulp at any magnitude for exponents of a few units (the error is
// that of log2 of the mantissa times v, independent of the base's exponent):
//   x = m 2^e with m in [sqrt(1/2), sqrt(2));  log2 m = (2 / ln 2) atanh((m - 1) / (m + 1)), odd series to s^9
//   (|s| <= 0.172, truncation 2e-9 relative);  v e is split exactly (product + FMA residual), so the integer
//   part of v log2 x never costs mantissa bits;  2^fraction on the SFU, 2^integer exact.
__device__ __forceinline__ float pow_core(int bits, int e0, float v)
{
    const int e = (bits - 0x3f3504f3) >> 23;
    const float m = __int_as_float(bits - (e << 23)), fe = (float)(e + e0);
    const float num = __fsub_rn(m, 1.0f), den = __fadd_rn(m, 1.0f);
    const float r = rcp_approx(den);
    float s = __fmul_rn(num, r);
    s = fmaf(r, fmaf(-s, den, num), s); // one Newton step: the quotient to ~0.5 ulp
    const float z = __fmul_rn(s, s);
    float p = fmaf(z, 1.0f / 9.0f, 1.0f / 7.0f);
    p = fmaf(p, z, 1.0f / 5.0f);
    p = fmaf(p, z, 1.0f / 3.0f);
    p = __fmul_rn(p, z);
    const float u = __fmul_rn(s, 2.8853900817779268f);
    const float lm = fmaf(u, p, u); // log2(m)
    const float a = __fmul_rn(v, fe);
    const float b = fmaf(lm, v, fmaf(fe, v, -a));
    // nearest integer by the 1.5 * 2^23 trick (|v log2 x| < 2^22 or the result is 0 / infinity anyway)
    const float t = fminf(fmaxf(__fadd_rn(a, b), -1000.0f), 1000.0f);
    const float n = __fsub_rn(__fadd_rn(t, 12582912.0f), 12582912.0f);
    const float f = __fadd_rn(__fsub_rn(a, n), b);
    // 2^n on the SFU: exact for integers; below 2^-126 it flushes to zero (1e-38 absolute)
    return __fmul_rn(ex2_approx(f), ex2_approx(n));
}
// The same with log2 of the mantissa on the SFU (lg2.approx: absolute error 2^-22.6 on [sqrt(1/2), sqrt(2))): 14 instructions
// fewer per value.  The result's relative error is |v| ln2 1.6e-7 + 2^-22 (the final ex2): below 1e-6 for |v| <= 4, which is what
// the callers check (every colour curve: exponents 1/2.4 .. 2.6) -- larger exponents keep the series above.
__device__ __forceinline__ float pow_core_sfu(int bits, int e0, float v)
{
    const int e = (bits - 0x3f3504f3) >> 23;
    const float m = __int_as_float(bits - (e << 23)), fe = (float)(e + e0);
    float lm;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lm) : "f"(m));
    const float a = __fmul_rn(v, fe);
    const float b = fmaf(lm, v, fmaf(fe, v, -a));
    const float t = fminf(fmaxf(__fadd_rn(a, b), -1000.0f), 1000.0f);
    const float n = __fsub_rn(__fadd_rn(t, 12582912.0f), 12582912.0f);
    const float f = __fadd_rn(__fsub_rn(a, n), b);
    return __fmul_rn(ex2_approx(f), ex2_approx(n));
}
// the rare bases: denormals are renormalised; zero, infinity and NaN take exp2f(v log2f(x))
__device__ __noinline__ float pow_rare(float x, float v)
{
    const int bits = __float_as_int(x);
    if (bits > 0 && bits < 0x00800000) return pow_core(__float_as_int(__fmul_rn(x, 16777216.0f)), -24, v);
    return exp2f(__fmul_rn(v, log2f(x)));
}
// x >= 0 (the colour curves)
__device__ __forceinline__ float pow_pos(float x, float v)
{
    const int bits = __float_as_int(x);
    if ((unsigned)(bits - 0x00800000) >= 0x7f000000u) return pow_rare(x, v);
    return fabsf(v) <= 4.0f ? pow_core_sfu(bits, 0, v) : pow_core(bits, 0, v);
}
// Everything that is not a positive normal number, with powf's rules for the sign bit: an odd integer
// exponent keeps it; a non-integer exponent of a negative finite non-zero base is NaN; -0 and -infinity
// behave as their magnitudes otherwise.  flags: bit 0 = v is an integer, bit 1 = v is an odd integer.
__device__ __noinline__ float pow_signed_rare(float x, float v, int flags)
{
    const float ax = fabsf(x);
    const float r = pow_pos(ax, v);
    if (__float_as_int(x) < 0) {
        if (flags & 2) return -r;
        if (!(flags & 1) && ax != 0.0f && ax <= 3.4028234664e38f) return __int_as_float(0x7fffffff);
    }
    return r;
}
// any base: one test sends positive normal numbers (every pixel of an ordinary image) to the core
__device__ __forceinline__ float pow_fast(float x, float v, int flags)
{
    const int bits = __float_as_int(x);
    if ((unsigned)(bits - 0x00800000) >= 0x7f000000u) return pow_signed_rare(x, v, flags);
    return fabsf(v) <= 4.0f ? pow_core_sfu(bits, 0, v) : pow_core(bits, 0, v);
}
struct PowOp {
    Src s;
    float v;
    TC_OP_1(s)
    __device__ __forceinline__ float4 apply(float4 p) const
    {
        // small integer exponents (the fixtures use 2 and 3) as products: within 1 ulp of powf
        if (v == 2.0f) return make_float4(p.x * p.x, p.y * p.y, p.z * p.z, p.w * p.w);
        if (v == 3.0f) return make_float4(p.x * p.x * p.x, p.y * p.y * p.y, p.z * p.z * p.z, p.w * p.w * p.w);
        if (v == 0.0f) return make_float4(1.f, 1.f, 1.f, 1.f);
        // any other exponent: exp2(v * log2|x|) with powf's sign rules (pow_fast, tolerance path)
        const bool v_int = truncf(v) == v, v_odd = v_int && fabsf(v) < 16777216.0f && ((long long)v & 1);
        const int flags = (v_int ? 1 : 0) | (v_odd ? 2 : 0);
        return make_float4(pow_fast(p.x, v, flags), pow_fast(p.y, v, flags), pow_fast(p.z, v, flags), pow_fast(p.w, v, flags));
    }
};
struct SaturateOp {
    Src s;
    float v;
    TC_OP_1(s)
    __device__ __forceinline__ float4 apply(float4 p) const
    {
        const float l = luma709(p);
        p.x = lerp_oiio(l, p.x, v);
        p.y = lerp_oiio(l, p.y, v);
        p.z = lerp_oiio(l, p.z, v);
        return p;
    }
};
struct ColorMapOp {
    Src s;
    int n;          // knots (0 = unknown map: RGB stays zero, alpha copied)
    const float* k; // n x RGB knots in device memory: neighbouring pixels look up different segments,
                    // which a constant-bank operand would serialise; L1-cached global loads do not
    TC_OP_1(s)
    __device__ __forceinline__ float4 apply(const float4 p) const
    {
        float4 o = make_float4(0.f, 0.f, 0.f, p.w);
        if (n) {
            float t = fminf(fmaxf(luma709(p), 0.0f), 1.0f);
            const int nsegs = n - 1;
            const float xs = __fmul_rn(t, (float)nsegs);
            const float fl = floorf(xs);
            const int seg = (int)fl;
            const float fr = __fsub_rn(xs, fl);
            const int nxt = min(seg + 1, nsegs);
            o.x = lerp_oiio(__ldg(k + seg * 3 + 0), __ldg(k + nxt * 3 + 0), fr);
            o.y = lerp_oiio(__ldg(k + seg * 3 + 1), __ldg(k + nxt * 3 + 1), fr);
            o.z = lerp_oiio(__ldg(k + seg * 3 + 2), __ldg(k + nxt * 3 + 2), fr);
        }
        return o;
    }
};
struct PremultOp {
    Src s;
    TC_OP_1(s)
    __device__ __forceinline__ float4 apply(float4 p) const
    {
        if (p.w != 1.0f) {
            p.x = __fmul_rn(p.x, p.w);
            p.y = __fmul_rn(p.y, p.w);
            p.z = __fmul_rn(p.z, p.w);
        }
        return p;
    }
};
struct UnpremultOp {
    Src s;
    TC_OP_1(s)
    __device__ __forceinline__ float4 apply(float4 p) const
    {
        if (p.w != 0.0f && p.w != 1.0f) {
            p.x = __fdiv_rn(p.x, p.w);
            p.y = __fdiv_rn(p.y, p.w);
            p.z = __fdiv_rn(p.z, p.w);
        }
        return p;
    }
};

// colour convert: decode curve -> 3x3 -> encode curve (alpha untouched)
__device__ __forceinline__ float cc_curve(int id, int encode, float p, float x)
{
    switch (id) {
    case 1:
        if (x <= 0.0f) return x;
        return pow_pos(x, encode ? __fdiv_rn(1.0f, p) : p);
    case 2:
        if (!encode) return x <= 0.04045f ? __fdiv_rn(x, 12.92f) : pow_pos(__fdiv_rn(__fadd_rn(x, 0.055f), 1.055f), 2.4f);
        return x <= 0.0031308f ? __fmul_rn(x, 12.92f) : __fsub_rn(__fmul_rn(1.055f, pow_pos(x, 1.0f / 2.4f)), 0.055f);
    case 3:
        if (encode)
            return x <= 0.0078125f ? __fadd_rn(__fmul_rn(10.5402377416545f, x), 0.0729055341958355f)
                                   : __fdiv_rn(__fadd_rn(log2f(x), 9.72f), 17.52f);
        return x <= 0.155251141552511f ? __fdiv_rn(__fsub_rn(x, 0.0729055341958355f), 10.5402377416545f)
                                       : exp2f(__fsub_rn(__fmul_rn(x, 17.52f), 9.72f));
    case 4:
        if (encode) {
            if (x <= 0.0f) return -0.3584474886f;
            if (x < 0.000030517578125f)
                return __fdiv_rn(__fadd_rn(log2f(__fadd_rn(0.0000152587890625f, __fmul_rn(x, 0.5f))), 9.72f), 17.52f);
            return __fdiv_rn(__fadd_rn(log2f(x), 9.72f), 17.52f);
        }
        if (x <= -0.3013698630f)
            return __fmul_rn(__fsub_rn(exp2f(__fsub_rn(__fmul_rn(x, 17.52f), 9.72f)), 0.0000152587890625f), 2.0f);
        return exp2f(__fsub_rn(__fmul_rn(x, 17.52f), 9.72f));
    default: return x;
    }
}
struct ColorConvertOp {
    Src s;
    int curve_in, curve_out;
    float p_in, p_out;
    float m[9];
    int unpremult;
    TC_OP_1(s)
    __device__ __forceinline__ float4 apply(float4 p) const
    {
        const bool un = unpremult && p.w != 0.0f && p.w != 1.0f;
        if (un) {
            p.x = __fdiv_rn(p.x, p.w);
            p.y = __fdiv_rn(p.y, p.w);
            p.z = __fdiv_rn(p.z, p.w);
        }
        const float l0 = cc_curve(curve_in, 0, p_in, p.x), l1 = cc_curve(curve_in, 0, p_in, p.y),
                    l2 = cc_curve(curve_in, 0, p_in, p.z);
        float o[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float v = __fadd_rn(__fadd_rn(__fmul_rn(m[c * 3 + 0], l0), __fmul_rn(m[c * 3 + 1], l1)),
                                      __fmul_rn(m[c * 3 + 2], l2));
            o[c] = cc_curve(curve_out, 1, p_out, v);
            if (un) o[c] = __fmul_rn(o[c], p.w);
        }
        return make_float4(o[0], o[1], o[2], p.w);
    }
};

// The same conversion with both curves fixed at compile time (the per-node kernel).  With run-time curve
// ids every pixel of the unrolled loop carries all five curves twice over and the kernel outgrows the
// instruction cache; a (decode, encode) pair is a few hundred instructions.  The chain kernel keeps the
// run-time form above.
template <int CIN, int COUT>
struct ColorConvertOpT {
    Src s;
    float p_in, p_out;
    float m[9];
    int unpremult;
    TC_OP_1(s)
    __device__ __forceinline__ float4 apply(float4 p) const
    {
        const bool un = unpremult && p.w != 0.0f && p.w != 1.0f;
        if (un) {
            p.x = __fdiv_rn(p.x, p.w);
            p.y = __fdiv_rn(p.y, p.w);
            p.z = __fdiv_rn(p.z, p.w);
        }
        const float l0 = cc_curve(CIN, 0, p_in, p.x), l1 = cc_curve(CIN, 0, p_in, p.y), l2 = cc_curve(CIN, 0, p_in, p.z);
        float o[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float v = __fadd_rn(__fadd_rn(__fmul_rn(m[c * 3 + 0], l0), __fmul_rn(m[c * 3 + 1], l1)),
                                      __fmul_rn(m[c * 3 + 2], l2));
            o[c] = cc_curve(COUT, 1, p_out, v);
            if (un) o[c] = __fmul_rn(o[c], p.w);
        }
        return make_float4(o[0], o[1], o[2], p.w);
    }
};
constexpr int CC_CURVES = 5; // linear, gamma, sRGB, ACEScct, ACEScc (cc_curve)

// ---- chain of pointwise effects in one pass -----------------------------------------------
// Each step is the per-node functor above, applied in registers; between steps the value takes the
// round trip through the images' storage type that the node boundary would have applied.
struct ChainStep {
    int kind;
    float v;          // Pow / Saturate value
    int n;            // ColorMap knots
    const float* k;   // ColorMap knot table (device)
    int curve_in, curve_out, unpremult;
    float p_in, p_out;
    float m[9];
};
// The steps of a chain applied to one pixel (shared by the plain and the flipped-source kernels).
struct ChainSteps {
    int n, type;
    ChainStep step[TC_CHAIN_MAX];
    __device__ __forceinline__ float4 apply(float4 p) const
    {
        for (int i = 0; i < n; ++i) {
            const ChainStep& st = step[i];
            switch (st.kind) {
            case TC_CHAIN_INVERT: p = InvertOp().apply(p); break;
            case TC_CHAIN_POW: {
                PowOp o;
                o.v = st.v;
                p = o.apply(p);
                break;
            }
            case TC_CHAIN_SATURATE: {
                SaturateOp o;
                o.v = st.v;
                p = o.apply(p);
                break;
            }
            case TC_CHAIN_COLOR_MAP: {
                ColorMapOp o;
                o.n = st.n;
                o.k = st.k;
                p = o.apply(p);
                break;
            }
            case TC_CHAIN_PREMULT: p = PremultOp().apply(p); break;
            case TC_CHAIN_UNPREMULT: p = UnpremultOp().apply(p); break;
            case TC_CHAIN_COLOR_CONVERT: {
                ColorConvertOp o;
                o.curve_in = st.curve_in, o.curve_out = st.curve_out, o.unpremult = st.unpremult;
                o.p_in = st.p_in, o.p_out = st.p_out;
#pragma unroll
                for (int c = 0; c < 9; ++c) o.m[c] = st.m[c];
                p = o.apply(p);
                break;
            }
            case TC_CHAIN_OVER_FILL: {
                // CompNode: premult(fg) [a typed buffer in the reference], then over the Fill frame's colour (m[0..3])
                if (st.unpremult) { // (the step's flag: premultiply first)
                    if (p.w != 1.0f) {
                        p.x = __fmul_rn(p.x, p.w);
                        p.y = __fmul_rn(p.y, p.w);
                        p.z = __fmul_rn(p.z, p.w);
                    }
                    p = quant4(p, type);
                }
                const float alpha = fminf(fmaxf(p.w, 0.0f), 1.0f);
                const float oma = __fsub_rn(1.0f, alpha);
                p = make_float4(__fadd_rn(p.x, __fmul_rn(oma, st.m[0])), __fadd_rn(p.y, __fmul_rn(oma, st.m[1])),
                                __fadd_rn(p.z, __fmul_rn(oma, st.m[2])), __fadd_rn(p.w, __fmul_rn(oma, st.m[3])));
                break;
            }
            default: p = make_float4(0.f, 0.f, 0.f, 0.f); break;
            }
            if (i + 1 < n) p = quant4(p, type);
        }
        return p;
    }
};
// The fast chain kernel: the step loop runs OUTSIDE the pixel loop -- a thread owns 4 (float / half) or 8 (8-bit) pixels and every
// step is dispatched once for all of them -- and the storage type is a template parameter, so the node-boundary quantisation is
// straight-line code.  (The generic k_pointwise<ChainOp> dispatches per pixel: 216 instructions per pixel for a three-step chain
// on half frames, half of them branches and constant loads; profiles/r02_chain_kernel.md.)  Same functors, same arithmetic, same
// quantisation points: bit-identical results.  Flip / Flop nodes read the mirrored 16-byte group and reverse it in registers.
template <int TYPE>
struct ChainGeom {
    static constexpr int VEC = TYPE == TC_F32 ? 1 : (TYPE == TC_U8 ? 4 : 2); // pixels per 16 bytes
    static constexpr int NG = TYPE == TC_F32 ? 4 : 2;                        // 16-byte groups per thread-iteration
    static constexpr int NP = VEC * NG;
};
template <int TYPE>
__device__ __forceinline__ void chain_unpack(const uint4 v, float4* px, bool reverse)
{
    constexpr int VEC = ChainGeom<TYPE>::VEC;
    float4 t[VEC];
    if (TYPE == TC_F32) {
        t[0] = make_float4(__uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z), __uint_as_float(v.w));
    } else if (TYPE == TC_U8) {
        const unsigned w[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
        for (int j = 0; j < 4; ++j) t[j] = make_float4(u8_to_f(w[j] & 0xffu), u8_to_f((w[j] >> 8) & 0xffu), u8_to_f((w[j] >> 16) & 0xffu), u8_to_f(w[j] >> 24));
    } else if (TYPE == TC_U16) {
        t[0] = make_float4(u16_to_f(v.x & 0xffffu), u16_to_f(v.x >> 16), u16_to_f(v.y & 0xffffu), u16_to_f(v.y >> 16));
        t[VEC - 1] = make_float4(u16_to_f(v.z & 0xffffu), u16_to_f(v.z >> 16), u16_to_f(v.w & 0xffffu), u16_to_f(v.w >> 16));
    } else {
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&v.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
        const float2 c = __half22float2(*reinterpret_cast<const __half2*>(&v.z)), d = __half22float2(*reinterpret_cast<const __half2*>(&v.w));
        t[0] = make_float4(a.x, a.y, b.x, b.y);
        t[VEC - 1] = make_float4(c.x, c.y, d.x, d.y);
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) px[j] = reverse ? t[VEC - 1 - j] : t[j];
}

// the two long bodies as calls (one copy of the code, the pixel array stays in registers: a rolled loop would index it dynamically)
__device__ __noinline__ float4 chain_pow_general(float v, float4 p)
{
    PowOp o;
    o.v = v;
    return o.apply(p);
}
__device__ __noinline__ float4 chain_color_convert(const ChainStep& st, float4 p)
{
    ColorConvertOp o;
    o.curve_in = st.curve_in, o.curve_out = st.curve_out, o.unpremult = st.unpremult;
    o.p_in = st.p_in, o.p_out = st.p_out;
#pragma unroll
    for (int m = 0; m < 9; ++m) o.m[m] = st.m[m];
    return o.apply(p);
}

template <int TYPE>
__global__ void __launch_bounds__(PW_THREADS) k_chain(const __grid_constant__ ChainSteps c, const Img src, const Img dst, const int flipx, const int flipy)
{
    using G = ChainGeom<TYPE>;
    constexpr int VEC = G::VEC, NG = G::NG, NP = G::NP;
    const unsigned groups = ((unsigned)dst.w * (unsigned)dst.h) / VEC;
    const unsigned gw = (unsigned)dst.w / VEC; // groups per row (flips only: the width is a multiple of VEC then)
    const unsigned stride = gridDim.x * blockDim.x;
    const uint4* __restrict__ in = reinterpret_cast<const uint4*>(src.p);
    for (unsigned base = blockIdx.x * blockDim.x + threadIdx.x; base < groups; base += NG * stride) {
        uint4 raw[NG];
#pragma unroll
        for (int k = 0; k < NG; ++k) {
            const unsigned g = base + k * stride;
            if (g < groups) {
                unsigned sg = g;
                if (flipx | flipy) {
                    const unsigned y = g / gw, xg = g - y * gw;
                    sg = (flipy ? (unsigned)dst.h - 1 - y : y) * gw + (flipx ? gw - 1 - xg : xg);
                }
                raw[k] = __ldg(in + sg);
            } else {
                raw[k] = make_uint4(0u, 0u, 0u, 0u);
            }
        }
        float4 px[NP];
#pragma unroll
        for (int k = 0; k < NG; ++k) chain_unpack<TYPE>(raw[k], px + k * VEC, flipx != 0);
        for (int i = 0; i < c.n; ++i) {
            const ChainStep& st = c.step[i];
            switch (st.kind) {
            case TC_CHAIN_INVERT:
#pragma unroll
                for (int q = 0; q < NP; ++q) px[q] = InvertOp().apply(px[q]);
                break;
            case TC_CHAIN_POW: {
                // (the exponent classes of PowOp::apply, decided once per step instead of once per pixel)
                if (st.v == 2.0f) {
#pragma unroll
                    for (int q = 0; q < NP; ++q) px[q] = make_float4(px[q].x * px[q].x, px[q].y * px[q].y, px[q].z * px[q].z, px[q].w * px[q].w);
                } else if (st.v == 3.0f) {
#pragma unroll
                    for (int q = 0; q < NP; ++q)
                        px[q] = make_float4(px[q].x * px[q].x * px[q].x, px[q].y * px[q].y * px[q].y, px[q].z * px[q].z * px[q].z, px[q].w * px[q].w * px[q].w);
                } else {
#pragma unroll
                    for (int q = 0; q < NP; ++q) px[q] = chain_pow_general(st.v, px[q]);
                }
                break;
            }
            case TC_CHAIN_SATURATE: {
                SaturateOp o;
                o.v = st.v;
#pragma unroll
                for (int q = 0; q < NP; ++q) px[q] = o.apply(px[q]);
                break;
            }
            case TC_CHAIN_COLOR_MAP: {
                ColorMapOp o;
                o.n = st.n;
                o.k = st.k;
#pragma unroll
                for (int q = 0; q < NP; ++q) px[q] = o.apply(px[q]);
                break;
            }
            case TC_CHAIN_PREMULT:
#pragma unroll
                for (int q = 0; q < NP; ++q) px[q] = PremultOp().apply(px[q]);
                break;
            case TC_CHAIN_UNPREMULT:
#pragma unroll
                for (int q = 0; q < NP; ++q) px[q] = UnpremultOp().apply(px[q]);
                break;
            case TC_CHAIN_COLOR_CONVERT:
#pragma unroll
                for (int q = 0; q < NP; ++q) px[q] = chain_color_convert(st, px[q]);
                break;
            case TC_CHAIN_OVER_FILL: {
                const float4 b = make_float4(st.m[0], st.m[1], st.m[2], st.m[3]);
#pragma unroll
                for (int q = 0; q < NP; ++q) {
                    float4 p = px[q];
                    if (st.unpremult) { // (the step's flag: premultiply first, into a typed buffer)
                        if (p.w != 1.0f) {
                            p.x = __fmul_rn(p.x, p.w);
                            p.y = __fmul_rn(p.y, p.w);
                            p.z = __fmul_rn(p.z, p.w);
                        }
                        p = quant4(p, TYPE);
                    }
                    const float oma = __fsub_rn(1.0f, fminf(fmaxf(p.w, 0.0f), 1.0f));
                    px[q] = make_float4(__fadd_rn(p.x, __fmul_rn(oma, b.x)), __fadd_rn(p.y, __fmul_rn(oma, b.y)), __fadd_rn(p.z, __fmul_rn(oma, b.z)),
                                        __fadd_rn(p.w, __fmul_rn(oma, b.w)));
                }
                break;
            }
            default:
#pragma unroll
                for (int q = 0; q < NP; ++q) px[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                break;
            }
            if (i + 1 < c.n && TYPE != TC_F32) {
#pragma unroll
                for (int q = 0; q < NP; ++q) px[q] = quant4(px[q], TYPE);
            }
        }
#pragma unroll
        for (int k = 0; k < NG; ++k) {
            const unsigned g = base + k * stride;
            if (g < groups) {
                float4 v[VEC];
#pragma unroll
                for (int j = 0; j < VEC; ++j) v[j] = px[k * VEC + j];
                store_vec<VEC>(dst.p, TYPE, g * VEC, v);
            }
        }
    }
}

template <int TYPE>
static int launch_chain(const ChainSteps& c, const tc_image* dst, const tc_image* src, int flipx, int flipy, tc_stream s)
{
    using G = ChainGeom<TYPE>;
    const size_t groups = (size_t)dst->w * dst->h / G::VEC;
    const size_t per_block = (size_t)PW_THREADS * G::NG;
    size_t blocks = (groups + per_block - 1) / per_block;
    const size_t cap = (size_t)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    k_chain<TYPE><<<(unsigned)blocks, PW_THREADS, 0, (cudaStream_t)s>>>(c, view(src), view(dst), flipx, flipy);
    count_launch();
    return check_cuda(cudaGetLastError(), "k_chain launch");
}

struct ChainOp {
    Src s;
    ChainSteps c;
    TC_OP_1(s)
    __device__ __forceinline__ float4 apply(float4 p) const { return c.apply(p); }
};
// ... with Flip / Flop nodes in the chain: they commute with every pointwise step, so their net effect is where the source is read
struct ChainFlipOp {
    TC_OP_0()
    Img src;
    int flipx, flipy;
    ChainSteps c;
    __device__ float4 operator()(int x, int y, unsigned) const
    {
        return c.apply(load_black(src, flipx ? src.w - 1 - x : x, flipy ? src.h - 1 - y : y));
    }
};

// ---- address-only transforms ---------------------------------------------------------
struct CropOp {
    TC_OP_0()
    Img src;
    int px, py, cw, ch;
    __device__ float4 operator()(int x, int y, unsigned) const
    {
        if (x >= cw || y >= ch) return make_float4(0.f, 0.f, 0.f, 0.f);
        return load_black(src, px + x, py + y);
    }
};
struct FlipOp {
    TC_OP_0()
    Img src;
    __device__ float4 operator()(int x, int y, unsigned) const { return load_black(src, x, src.h - 1 - y); }
};
struct FlopOp {
    TC_OP_0()
    Img src;
    __device__ float4 operator()(int x, int y, unsigned) const { return load_black(src, src.w - 1 - x, y); }
};
struct ConvertOp {
    Src s;
    TC_OP_1(s)
    __device__ __forceinline__ float4 apply(float4 p) const { return p; }
};
struct PasteOp {
    TC_OP_0()
    Img src;
    int ox, oy;
    __device__ float4 operator()(int x, int y, unsigned) const { return load_black(src, x - ox, y - oy); }
};

// ---- transitions -----------------------------------------------------------------------
struct DissolveOp {
    Src a, b;
    float qiv, qv; // the fill constants, already quantised to each source's storage type
    TC_OP_2(a, b)
    __device__ __forceinline__ float4 apply(const float4 p, const float4 q, int, int) const
    {
        const int ta = a.im.type, tb = b.im.type;
        return make_float4(__fadd_rn(quant(__fmul_rn(p.x, qiv), ta), quant(__fmul_rn(q.x, qv), tb)),
                           __fadd_rn(quant(__fmul_rn(p.y, qiv), ta), quant(__fmul_rn(q.y, qv), tb)),
                           __fadd_rn(quant(__fmul_rn(p.z, qiv), ta), quant(__fmul_rn(q.z, qv), tb)),
                           __fadd_rn(quant(__fmul_rn(p.w, qiv), ta), quant(__fmul_rn(q.w, qv), tb)));
    }
};
struct WipeOp {
    Src a, b;
    int mw, mh; // matte size (= `from` size)
    int edge, vertical;
    TC_OP_2(a, b)
    __device__ __forceinline__ float4 apply(const float4 p, const float4 q, int x, int y) const
    {
        float mf = 0.f, mt = 0.f;
        if (x < mw && y < mh) {
            const int k = (vertical ? y : x) - edge;
            if (k < 0) {
                mf = 0.f;
                mt = 1.f;
            } else if (k < 200) {
                // 200-row UINT8 ramp: row r = u8(lerp(0, 1, r/199))
                mf = quant(lerp_oiio(0.f, 1.f, __fdiv_rn((float)k, 199.0f)), TC_U8);
                mt = quant(lerp_oiio(0.f, 1.f, __fdiv_rn((float)(199 - k), 199.0f)), TC_U8);
            } else {
                mf = 1.f;
                mt = 0.f;
            }
        }
        const int ta = a.im.type, tb = b.im.type;
        return make_float4(__fadd_rn(quant(__fmul_rn(p.x, mf), ta), quant(__fmul_rn(q.x, mt), tb)),
                           __fadd_rn(quant(__fmul_rn(p.y, mf), ta), quant(__fmul_rn(q.y, mt), tb)),
                           __fadd_rn(quant(__fmul_rn(p.z, mf), ta), quant(__fmul_rn(q.z, mt), tb)),
                           __fadd_rn(quant(__fmul_rn(p.w, mf), ta), quant(__fmul_rn(q.w, mt), tb)));
    }
};

// ---- compositor: premult(fg) (stored in fg's type) + over ---------------------------
struct OverOp {
    Src fg, bg;
    int premult;
    TC_OP_2(fg, bg)
    __device__ __forceinline__ float4 apply(float4 a, const float4 b, int, int) const
    {
        if (premult) {
            if (a.w != 1.0f) {
                a.x = __fmul_rn(a.x, a.w);
                a.y = __fmul_rn(a.y, a.w);
                a.z = __fmul_rn(a.z, a.w);
            }
            a = quant4(a, fg.im.type);
        }
        const float alpha = fminf(fmaxf(a.w, 0.0f), 1.0f);
        const float oma = __fsub_rn(1.0f, alpha);
        return make_float4(__fadd_rn(a.x, __fmul_rn(oma, b.x)), __fadd_rn(a.y, __fmul_rn(oma, b.y)),
                           __fadd_rn(a.z, __fmul_rn(oma, b.z)), __fadd_rn(a.w, __fmul_rn(oma, b.w)));
    }
};
// over a constant background: the Fill generator at the bottom of every track stack (ImageGraph.cpp:165-171)
// never needs to exist as a frame -- same arithmetic as OverOp with b = the fill colour as the Fill
// generator's storage type holds it
struct OverFillOp {
    Src fg;
    float4 b;
    int premult;
    TC_OP_1(fg)
    __device__ __forceinline__ float4 apply(float4 a) const
    {
        if (premult) {
            if (a.w != 1.0f) {
                a.x = __fmul_rn(a.x, a.w);
                a.y = __fmul_rn(a.y, a.w);
                a.z = __fmul_rn(a.z, a.w);
            }
            a = quant4(a, fg.im.type);
        }
        const float alpha = fminf(fmaxf(a.w, 0.0f), 1.0f);
        const float oma = __fsub_rn(1.0f, alpha);
        return make_float4(__fadd_rn(a.x, __fmul_rn(oma, b.x)), __fadd_rn(a.y, __fmul_rn(oma, b.y)),
                           __fadd_rn(a.z, __fmul_rn(oma, b.z)), __fadd_rn(a.w, __fmul_rn(oma, b.w)));
    }
};

// colour-map knot tables (tc_tables.cpp)
int color_map_knots(const char* name, const float** knots);
// colour-space table (tc_tables.cpp)
int color_convert_setup(const char* from, const char* to, int* curve_in, float* p_in, float m[9], int* curve_out, float* p_out);

// Knot tables resident in device memory, uploaded once per (device, table).
static int device_knots(const float* host_knots, int n, tc_stream s, const float** out)
{
    struct Entry {
        int device;
        const float* host;
        float* dev;
    };
    static std::mutex mutex;
    static std::vector<Entry> entries;
    int device = 0;
    TC_CUDA(cudaGetDevice(&device));
    std::lock_guard<std::mutex> lock(mutex);
    for (const Entry& e : entries)
        if (e.device == device && e.host == host_knots) {
            *out = e.dev;
            return TC_OK;
        }
    float* d = nullptr;
    TC_CUDA(cudaMalloc(&d, sizeof(float) * 3 * n));
    TC_CUDA(cudaMemcpy(d, host_knots, sizeof(float) * 3 * n, cudaMemcpyHostToDevice));
    entries.push_back(Entry { device, host_knots, d });
    *out = d;
    return TC_OK;
}

static inline float4 f4(const float c[4]) { return make_float4(c[0], c[1], c[2], c[3]); }

// host-side restatement of the storage round trip (for constants baked on the host)
static float quant_host(float v, int type)
{
    auto q = [](float f, float scale) {
        float s = f * scale;
        s += (s < 0.0f ? -0.5f : 0.5f);
        s = s < 0.0f ? 0.0f : (s > scale ? scale : s);
        return (float)(unsigned)s * (1.0f / scale);
    };
    switch (type) {
    case TC_U8: return q(v, 255.0f);
    case TC_U16: return q(v, 65535.0f);
    case TC_F16: return __half2float(__float2half_rn(v));
    default: return v;
    }
}

// raw output packing: RGBA working format -> `channels` x `type`, tightly packed
template <int CH>
__global__ void __launch_bounds__(256) k_pack_raw(void* __restrict__ dst, const Img src, int type)
{
    const unsigned total = (unsigned)src.w * (unsigned)src.h;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const float4 p = load_px(src.p, src.type, i);
        const float v[4] = { p.x, p.y, p.z, p.w };
        if (CH == 4) {
            store_px(dst, type, i, p);
        } else {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const size_t o = (size_t)i * 3 + c;
                switch (type) {
                case TC_U8: reinterpret_cast<unsigned char*>(dst)[o] = (unsigned char)f_to_u8(v[c]); break;
                case TC_U16: reinterpret_cast<unsigned short*>(dst)[o] = (unsigned short)f_to_u16(v[c]); break;
                case TC_F16: reinterpret_cast<__half*>(dst)[o] = __float2half_rn(v[c]); break;
                default: reinterpret_cast<float*>(dst)[o] = v[c]; break;
                }
            }
        }
    }
}

} // namespace tc

using namespace tc;

#define TC_VALIDATE(im, name)                   \
    do {                                        \
        int _v = validate(im, name);            \
        if (_v != TC_OK) return _v;             \
    } while (0)

extern "C" {

int tc_fill(const tc_image* dst, const float color[4], tc_stream s)
{
    TC_VALIDATE(dst, "tc_fill dst");
    if (!color) return fail(TC_ERR_INVALID, "tc_fill: null color");
    return launch(FillOp { f4(color) }, dst, s);
}

int tc_checkers(const tc_image* dst, int cw, int ch, const float c1[4], const float c2[4], tc_stream s)
{
    TC_VALIDATE(dst, "tc_checkers dst");
    if (!c1 || !c2) return fail(TC_ERR_INVALID, "tc_checkers: null color");
    if (cw <= 0 || ch <= 0) return fail(TC_ERR_INVALID, "tc_checkers: checker size %dx%d (the reference divides by zero here)", cw, ch);
    // exact division by multiplication: valid when (extent - 1) * checker < 2^32 and the checker is at least 2
    auto magic = [](int c, int extent) -> unsigned {
        if (c < 2 || (unsigned long long)(extent > 0 ? extent - 1 : 0) * (unsigned long long)c >= (1ull << 32)) return 0u;
        return (unsigned)((1ull << 32) / (unsigned)c) + 1u;
    };
    return launch(CheckerOp { f4(c1), f4(c2), cw, ch, magic(cw, dst->w), magic(ch, dst->h) }, dst, s);
}

int tc_gradient(const tc_image* dst, const float c1[4], const float c2[4], int vertical, tc_stream s)
{
    TC_VALIDATE(dst, "tc_gradient dst");
    if (!c1 || !c2) return fail(TC_ERR_INVALID, "tc_gradient: null color");
    const int n = (vertical ? dst->h : dst->w) - 1;
    return launch(GradientOp { f4(c1), f4(c2), (float)(n > 1 ? n : 1), vertical ? 1 : 0 }, dst, s);
}

int tc_noise(const tc_image* dst, const char* type, float a, float b, int mono, int seed, tc_stream s)
{
    TC_VALIDATE(dst, "tc_noise dst");
    int t = -1;
    if (type) {
        if (!strcmp(type, "gaussian") || !strcmp(type, "normal")) t = 0;
        else if (!strcmp(type, "white") || !strcmp(type, "uniform")) t = 1;
        else if (!strcmp(type, "salt")) t = 2;
    }
    if (t < 0) {
        // OIIO noise() fails on an unknown type and the zero-initialised buffer is returned
        const float z[4] = { 0.f, 0.f, 0.f, 0.f };
        return launch(FillOp { f4(z) }, dst, s);
    }
    return launch(NoiseOp { a, b, t, mono ? 1 : 0, seed }, dst, s);
}

int tc_invert(const tc_image* dst, const tc_image* src, tc_stream s)
{
    TC_VALIDATE(dst, "tc_invert dst");
    TC_VALIDATE(src, "tc_invert src");
    return launch(InvertOp { make_src(src, dst) }, dst, s);
}
int tc_pow(const tc_image* dst, const tc_image* src, float value, tc_stream s)
{
    TC_VALIDATE(dst, "tc_pow dst");
    TC_VALIDATE(src, "tc_pow src");
    return launch(PowOp { make_src(src, dst), value }, dst, s);
}
int tc_saturate(const tc_image* dst, const tc_image* src, float value, tc_stream s)
{
    TC_VALIDATE(dst, "tc_saturate dst");
    TC_VALIDATE(src, "tc_saturate src");
    return launch(SaturateOp { make_src(src, dst), value }, dst, s);
}
int tc_color_map(const tc_image* dst, const tc_image* src, const char* map_name, tc_stream s)
{
    TC_VALIDATE(dst, "tc_color_map dst");
    TC_VALIDATE(src, "tc_color_map src");
    ColorMapOp op;
    op.s = make_src(src, dst);
    const float* knots = nullptr;
    op.n = map_name ? color_map_knots(map_name, &knots) : 0;
    op.k = nullptr;
    if (op.n) {
        int st = device_knots(knots, op.n, s, &op.k);
        if (st != TC_OK) return st;
    }
    return launch(op, dst, s);
}
int tc_premult(const tc_image* dst, const tc_image* src, tc_stream s)
{
    TC_VALIDATE(dst, "tc_premult dst");
    TC_VALIDATE(src, "tc_premult src");
    return launch(PremultOp { make_src(src, dst) }, dst, s);
}
int tc_unpremult(const tc_image* dst, const tc_image* src, tc_stream s)
{
    TC_VALIDATE(dst, "tc_unpremult dst");
    TC_VALIDATE(src, "tc_unpremult src");
    return launch(UnpremultOp { make_src(src, dst) }, dst, s);
}
int tc_color_convert(const tc_image* dst, const tc_image* src, const char* fromspace, const char* tospace, tc_stream s)
{
    return tc_color_convert2(dst, src, fromspace, tospace, 0, s);
}
int tc_color_convert2(const tc_image* dst, const tc_image* src, const char* fromspace, const char* tospace, int unpremult, tc_stream s)
{
    TC_VALIDATE(dst, "tc_color_convert dst");
    TC_VALIDATE(src, "tc_color_convert src");
    ColorConvertOp op;
    op.s = make_src(src, dst);
    op.unpremult = unpremult ? 1 : 0;
    if (!fromspace || !tospace ||
        !color_convert_setup(fromspace, tospace, &op.curve_in, &op.p_in, op.m, &op.curve_out, &op.p_out)) {
        // OIIO colorconvert() fails on an unknown colour space; the output keeps its zero fill
        const float z[4] = { 0.f, 0.f, 0.f, 0.f };
        return launch(FillOp { f4(z) }, dst, s);
    }
    if (op.curve_in < 0 || op.curve_in >= CC_CURVES || op.curve_out < 0 || op.curve_out >= CC_CURVES) return launch(op, dst, s);
    switch (op.curve_in * CC_CURVES + op.curve_out) {
#define TC_CC(I, O)                                                    \
    case I * CC_CURVES + O: {                                          \
        ColorConvertOpT<I, O> t;                                       \
        t.s = op.s, t.p_in = op.p_in, t.p_out = op.p_out, t.unpremult = op.unpremult; \
        memcpy(t.m, op.m, sizeof(t.m));                                \
        return launch(t, dst, s);                                      \
    }
#define TC_CC_ROW(I) TC_CC(I, 0) TC_CC(I, 1) TC_CC(I, 2) TC_CC(I, 3) TC_CC(I, 4)
        TC_CC_ROW(0) TC_CC_ROW(1) TC_CC_ROW(2) TC_CC_ROW(3) TC_CC_ROW(4)
#undef TC_CC_ROW
#undef TC_CC
    default: return launch(op, dst, s);
    }
}

int tc_pointwise_chain(const tc_image* dst, const tc_image* src, const tc_chain_op* ops, int n, tc_stream s)
{
    TC_VALIDATE(dst, "tc_pointwise_chain dst");
    TC_VALIDATE(src, "tc_pointwise_chain src");
    if (!ops || n < 1 || n > TC_CHAIN_MAX) return fail(TC_ERR_INVALID, "tc_pointwise_chain: %d ops (1..%d)", n, TC_CHAIN_MAX);
    if (dst->type != src->type) return fail(TC_ERR_INVALID, "tc_pointwise_chain: source and destination storage types differ");
    if (TC_CHAIN_BLUR == ops[0].kind || TC_CHAIN_UNSHARP == ops[0].kind) {
        // A chain that starts with a convolution node.  Simple followers (Invert, Premult, Pow 2 / 3, over-Fill) run as the
        // convolution kernel's epilogue: one launch, no intermediate frame.  Anything else: the convolution into a temporary,
        // then the rest of the chain -- still one frame pass fewer per fused node.
        if (src->w != dst->w || src->h != dst->h) return fail(TC_ERR_INVALID, "tc_pointwise_chain: a convolution node needs same-sized frames");
        for (int i = 1; i < n; ++i)
            if (TC_CHAIN_BLUR == ops[i].kind || TC_CHAIN_UNSHARP == ops[i].kind)
                return fail(TC_ERR_INVALID, "tc_pointwise_chain: a convolution node can only start a chain");
        EpiSteps epi;
        memset(&epi, 0, sizeof(epi));
        epi.type = dst->type;
        bool simple = n - 1 <= 8;
        for (int i = 1; i < n && simple; ++i) {
            const tc_chain_op& o = ops[i];
            const int k = epi.n;
            switch (o.kind) {
            case TC_CHAIN_INVERT:
            case TC_CHAIN_PREMULT: break;
            case TC_CHAIN_POW: simple = o.value == 2.0f || o.value == 3.0f; break;
            case TC_CHAIN_OVER_FILL:
                if (o.w != dst->w || o.h != dst->h) return fail(TC_ERR_INVALID, "tc_pointwise_chain: the Fill frame's size differs from the chain's");
                epi.premult[k] = o.flag ? 1 : 0;
                for (int c = 0; c < 4; ++c) epi.color[k][c] = quant_host(o.color[c], TC_U8);
                break;
            default: simple = false; break;
            }
            epi.kind[k] = o.kind;
            epi.v[k] = o.value;
            ++epi.n;
        }
        if (simple && n > 1) {
            int fused = 0;
            const int e = gauss_with_epilogue(dst, src, ops[0], epi, s, &fused);
            if (e != TC_OK || fused) return e;
        }
        // not fusable into the kernel: convolution node, then the followers
        const tc_image* conv_dst = dst;
        tc_image tmp = *dst;
        if (n > 1) {
            int e = tc_malloc(&tmp.data, tc_image_bytes(dst->w, dst->h, dst->type), s);
            if (e != TC_OK) return e;
            conv_dst = &tmp;
        }
        int e = TC_CHAIN_BLUR == ops[0].kind ? tc_blur(conv_dst, src, ops[0].value, s)
                                             : tc_unsharp_mask(conv_dst, src, ops[0].name_a, ops[0].value, ops[0].color[0], ops[0].color[1], s);
        if (e == TC_OK && n > 1) e = tc_pointwise_chain(dst, &tmp, ops + 1, n - 1, s);
        if (n > 1) tc_free(tmp.data, s);
        return e;
    }
    ChainSteps steps;
    memset(&steps, 0, sizeof(steps));
    steps.type = dst->type;
    int flipx = 0, flipy = 0;
    for (int i = 0; i < n; ++i) {
        if (TC_CHAIN_FLIP == ops[i].kind || TC_CHAIN_FLOP == ops[i].kind) {
            // position-only nodes: their typed output equals their input, so no step (and no quantisation) remains of them
            if (src->w != dst->w || src->h != dst->h) return fail(TC_ERR_INVALID, "tc_pointwise_chain: Flip / Flop need same-sized frames");
            (TC_CHAIN_FLIP == ops[i].kind ? flipy : flipx) ^= 1;
            continue;
        }
        ChainStep& st = steps.step[steps.n++];
        st.kind = ops[i].kind;
        st.v = ops[i].value;
        switch (ops[i].kind) {
        case TC_CHAIN_OVER_FILL:
            if (ops[i].w != dst->w || ops[i].h != dst->h) return fail(TC_ERR_INVALID, "tc_pointwise_chain: the Fill frame's size differs from the chain's");
            st.unpremult = ops[i].flag ? 1 : 0;
            // the colour as the Fill generator's UINT8 frame holds it (ImageEffect.cpp:93)
            for (int c = 0; c < 4; ++c) st.m[c] = quant_host(ops[i].color[c], TC_U8);
            break;
        case TC_CHAIN_INVERT:
        case TC_CHAIN_POW:
        case TC_CHAIN_SATURATE:
        case TC_CHAIN_PREMULT:
        case TC_CHAIN_UNPREMULT:
        case TC_CHAIN_ZERO: break;
        case TC_CHAIN_COLOR_MAP: {
            const float* knots = nullptr;
            st.n = color_map_knots(ops[i].name_a, &knots);
            if (st.n) {
                int e = device_knots(knots, st.n, s, &st.k);
                if (e != TC_OK) return e;
            }
            break;
        }
        case TC_CHAIN_COLOR_CONVERT:
            st.unpremult = ops[i].flag ? 1 : 0;
            if (!color_convert_setup(ops[i].name_a, ops[i].name_b, &st.curve_in, &st.p_in, st.m, &st.curve_out, &st.p_out))
                st.kind = TC_CHAIN_ZERO; // colorconvert() fails: the node's output keeps its zero fill
            break;
        default: return fail(TC_ERR_INVALID, "tc_pointwise_chain: unknown op kind %d", ops[i].kind);
        }
    }
    // the fast kernel: whole 16-byte groups of same-sized, 16-byte aligned frames (a flip also needs whole groups per row)
    {
        const int vec = dst->type == TC_F32 ? 1 : (dst->type == TC_U8 ? 4 : 2);
        const size_t total = (size_t)dst->w * dst->h;
        if (steps.n > 0 && src->w == dst->w && src->h == dst->h && total < (1ull << 31) && total % vec == 0 && !(((uintptr_t)src->data | (uintptr_t)dst->data) & 15) &&
            (!(flipx || flipy) || dst->w % vec == 0)) {
            switch (dst->type) {
            case TC_U8: return launch_chain<TC_U8>(steps, dst, src, flipx, flipy, s);
            case TC_U16: return launch_chain<TC_U16>(steps, dst, src, flipx, flipy, s);
            case TC_F16: return launch_chain<TC_F16>(steps, dst, src, flipx, flipy, s);
            default: return launch_chain<TC_F32>(steps, dst, src, flipx, flipy, s);
            }
        }
    }
    if (flipx || flipy) {
        ChainFlipOp op;
        memset(&op, 0, sizeof(op));
        op.src = view(src);
        op.flipx = flipx;
        op.flipy = flipy;
        op.c = steps;
        if (0 == steps.n && !(flipx && flipy)) return flipy ? tc_flip(dst, src, s) : tc_flop(dst, src, s);
        return launch(op, dst, s); // (no steps: Flip + Flop, a copy read backwards along both axes)
    }
    // nodes whose mirrorings cancel (Flip, Flip) and nothing else: the copy each of them makes
    if (0 == steps.n) return tc_convert(dst, src, s);
    ChainOp op;
    memset(&op, 0, sizeof(op));
    op.s = make_src(src, dst);
    op.c = steps;
    return launch(op, dst, s);
}

int tc_crop(const tc_image* dst, const tc_image* src, int px, int py, int cw, int ch, tc_stream s)
{
    TC_VALIDATE(dst, "tc_crop dst");
    TC_VALIDATE(src, "tc_crop src");
    return launch(CropOp { view(src), px, py, cw, ch }, dst, s);
}
int tc_flip(const tc_image* dst, const tc_image* src, tc_stream s)
{
    TC_VALIDATE(dst, "tc_flip dst");
    TC_VALIDATE(src, "tc_flip src");
    return launch(FlipOp { view(src) }, dst, s);
}
int tc_flop(const tc_image* dst, const tc_image* src, tc_stream s)
{
    TC_VALIDATE(dst, "tc_flop dst");
    TC_VALIDATE(src, "tc_flop src");
    return launch(FlopOp { view(src) }, dst, s);
}
int tc_convert(const tc_image* dst, const tc_image* src, tc_stream s)
{
    TC_VALIDATE(dst, "tc_convert dst");
    TC_VALIDATE(src, "tc_convert src");
    return launch(ConvertOp { make_src(src, dst) }, dst, s);
}
int tc_paste_on_black(const tc_image* dst, int ox, int oy, const tc_image* src, tc_stream s)
{
    TC_VALIDATE(dst, "tc_paste_on_black dst");
    TC_VALIDATE(src, "tc_paste_on_black src");
    return launch(PasteOp { view(src), ox, oy }, dst, s);
}

int tc_pack_raw(void* dst_dev, const tc_image* src, int type, int channels, tc_stream s)
{
    TC_VALIDATE(src, "tc_pack_raw src");
    if (!dst_dev) return fail(TC_ERR_INVALID, "tc_pack_raw: null destination");
    if (type < TC_U8 || type > TC_F32 || (channels != 3 && channels != 4)) return fail(TC_ERR_INVALID, "tc_pack_raw: %d x type %d", channels, type);
    const size_t total = (size_t)src->w * src->h;
    size_t blocks = (total + 255) / 256;
    const size_t cap = (size_t)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    if (channels == 4) k_pack_raw<4><<<(unsigned)blocks, 256, 0, (cudaStream_t)s>>>(dst_dev, view(src), type);
    else k_pack_raw<3><<<(unsigned)blocks, 256, 0, (cudaStream_t)s>>>(dst_dev, view(src), type);
    count_launch();
    return check_cuda(cudaGetLastError(), "k_pack_raw launch");
}

int tc_dissolve(const tc_image* dst, const tc_image* from, const tc_image* to, double value, tc_stream s)
{
    TC_VALIDATE(dst, "tc_dissolve dst");
    TC_VALIDATE(from, "tc_dissolve from");
    TC_VALIDATE(to, "tc_dissolve to");
    const float v = (float)value;
    const float iv = (float)(1.0 - value);
    return launch(DissolveOp { make_src(from, dst), make_src(to, dst), quant_host(iv, from->type), quant_host(v, to->type) },
                  dst, s);
}
int tc_wipe(const tc_image* dst, const tc_image* from, const tc_image* to, double value, int vertical, tc_stream s)
{
    TC_VALIDATE(dst, "tc_wipe dst");
    TC_VALIDATE(from, "tc_wipe from");
    TC_VALIDATE(to, "tc_wipe to");
    const int edge = vertical ? (int)(from->h * value) : (int)(from->w * value);
    return launch(WipeOp { make_src(from, dst), make_src(to, dst), from->w, from->h, edge, vertical ? 1 : 0 }, dst, s);
}
int tc_over(const tc_image* dst, const tc_image* fg, const tc_image* bg, int premult_fg, tc_stream s)
{
    TC_VALIDATE(dst, "tc_over dst");
    TC_VALIDATE(fg, "tc_over fg");
    TC_VALIDATE(bg, "tc_over bg");
    return launch(OverOp { make_src(fg, dst), make_src(bg, dst), premult_fg ? 1 : 0 }, dst, s);
}
int tc_over_fill(const tc_image* dst, const tc_image* fg, const float color[4], int bg_type, int premult_fg, tc_stream s)
{
    TC_VALIDATE(dst, "tc_over_fill dst");
    TC_VALIDATE(fg, "tc_over_fill fg");
    if (!color || bg_type < TC_U8 || bg_type > TC_F32) return fail(TC_ERR_INVALID, "tc_over_fill: bad background");
    if (fg->w != dst->w || fg->h != dst->h) return fail(TC_ERR_INVALID, "tc_over_fill: foreground and destination sizes differ");
    float4 b;
    b.x = quant_host(color[0], bg_type);
    b.y = quant_host(color[1], bg_type);
    b.z = quant_host(color[2], bg_type);
    b.w = quant_host(color[3], bg_type);
    return launch(OverFillOp { make_src(fg, dst), b, premult_fg ? 1 : 0 }, dst, s);
}

} // extern "C"
