// tc_common.cuh -- shared device helpers: typed RGBA pixel load/store with OIIO's
// storage-conversion rules, error plumbing and launch helpers.  sm_100a only.
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/toucan_cuda.h"

namespace tc {

// ---- error plumbing (tc_api.cu) -------------------------------------------------------
int fail(int status, const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
void count_launch(unsigned n = 1);
int num_sms();
// Opt a kernel into `bytes` of dynamic shared memory (and the maximum shared-memory carve-out)
// on the CURRENT device.  Function attributes are per device, so this is tracked per
// (kernel, device): one process may drive all GPUs of a box (toucan::FrameScheduler).
int ensure_dynamic_smem(const void* kernel, size_t bytes);

#define TC_CUDA(call)                                      \
    do {                                                   \
        int _st = ::tc::check_cuda((call), #call);         \
        if (_st != TC_OK) return _st;                      \
    } while (0)

// Blur / UnsharpMask with an epilogue of simple chain steps (tc_conv.cu; called by tc_pointwise_chain for a chain that starts
// with a convolution node).  *fused = 0 when the case is not covered (nothing launched then).
struct EpiSteps {
    int n, type;
    int kind[8];     // TC_CHAIN_INVERT, _PREMULT, _POW (value 2 or 3 only), _OVER_FILL
    float v[8];      // Pow exponent
    int premult[8];  // OVER_FILL: premultiply first
    float color[8][4];
};
int gauss_with_epilogue(const tc_image* dst, const tc_image* src, const tc_chain_op& head, const EpiSteps& epi, tc_stream s, int* fused);

// Device-side image view.
struct Img {
    void* p;
    int w, h;
    int type;
};
static inline Img view(const tc_image* im) { return Img { im->data, im->w, im->h, im->type }; }
int validate(const tc_image* im, const char* name);

// ---- storage conversion: SURVEY 9.1 (OIIO convert_type) ------------------------------
__device__ __forceinline__ float u8_to_f(unsigned v) { return __fmul_rn((float)v, 1.0f / 255.0f); }
__device__ __forceinline__ float u16_to_f(unsigned v) { return __fmul_rn((float)v, 1.0f / 65535.0f); }
// OIIO's convert_type<float, unsigned>: clamp(f * max + (f < 0 ? -0.5 : 0.5), 0, max), then a C cast (truncation).
// A negative product ends at 0 either way (+0.5 leaves it below 1), and the saturating round-toward-zero convert
// (cvt.rzi.sat: NaN -> 0, below 0 -> 0, above max -> max) IS clamp-then-truncate -- three instructions instead of seven,
// which matters because the 8/16-bit pointwise kernels are bound by instruction issue, not by HBM.
__device__ __forceinline__ unsigned f_to_u8(float f)
{
    const float s = __fadd_rn(__fmul_rn(f, 255.0f), 0.5f);
    unsigned r;
    asm("cvt.rzi.sat.u8.f32 %0, %1;" : "=r"(r) : "f"(s));
    return r;
}
__device__ __forceinline__ unsigned f_to_u16(float f)
{
    const float s = __fadd_rn(__fmul_rn(f, 65535.0f), 0.5f);
    unsigned r;
    asm("cvt.rzi.sat.u16.f32 %0, %1;" : "=r"(r) : "f"(s));
    return r;
}

// float -> storage type -> float round trip (node-boundary quantisation)
__device__ __forceinline__ float quant(float v, int type)
{
    switch (type) {
    case TC_U8: return u8_to_f(f_to_u8(v));
    case TC_U16: return u16_to_f(f_to_u16(v));
    case TC_F16: return __half2float(__float2half_rn(v));
    default: return v;
    }
}
__device__ __forceinline__ float4 quant4(float4 v, int type)
{
    if (type == TC_F32) return v;
    return make_float4(quant(v.x, type), quant(v.y, type), quant(v.z, type), quant(v.w, type));
}

// Pixel load by linear pixel index.  One 32/64/128-bit transaction per pixel.
__device__ __forceinline__ float4 load_px(const void* __restrict__ base, int type, size_t i)
{
    switch (type) {
    case TC_U8: {
        uchar4 v = __ldg(reinterpret_cast<const uchar4*>(base) + i);
        return make_float4(u8_to_f(v.x), u8_to_f(v.y), u8_to_f(v.z), u8_to_f(v.w));
    }
    case TC_U16: {
        ushort4 v = __ldg(reinterpret_cast<const ushort4*>(base) + i);
        return make_float4(u16_to_f(v.x), u16_to_f(v.y), u16_to_f(v.z), u16_to_f(v.w));
    }
    case TC_F16: {
        uint2 raw = __ldg(reinterpret_cast<const uint2*>(base) + i);
        __half2 a = *reinterpret_cast<__half2*>(&raw.x), b = *reinterpret_cast<__half2*>(&raw.y);
        float2 fa = __half22float2(a), fb = __half22float2(b);
        return make_float4(fa.x, fa.y, fb.x, fb.y);
    }
    default: return __ldg(reinterpret_cast<const float4*>(base) + i);
    }
}
__device__ __forceinline__ void store_px(void* __restrict__ base, int type, size_t i, float4 v)
{
    switch (type) {
    case TC_U8:
        reinterpret_cast<uchar4*>(base)[i] = make_uchar4((unsigned char)f_to_u8(v.x), (unsigned char)f_to_u8(v.y),
                                                         (unsigned char)f_to_u8(v.z), (unsigned char)f_to_u8(v.w));
        break;
    case TC_U16:
        reinterpret_cast<ushort4*>(base)[i] = make_ushort4((unsigned short)f_to_u16(v.x), (unsigned short)f_to_u16(v.y),
                                                           (unsigned short)f_to_u16(v.z), (unsigned short)f_to_u16(v.w));
        break;
    case TC_F16: {
        __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
        uint2 raw;
        raw.x = *reinterpret_cast<unsigned*>(&a);
        raw.y = *reinterpret_cast<unsigned*>(&b);
        reinterpret_cast<uint2*>(base)[i] = raw;
        break;
    }
    default: reinterpret_cast<float4*>(base)[i] = v; break;
    }
}
// black outside the data window
__device__ __forceinline__ float4 load_black(const Img& im, int x, int y)
{
    if (x < 0 || y < 0 || x >= im.w || y >= im.h) return make_float4(0.f, 0.f, 0.f, 0.f);
    return load_px(im.p, im.type, (size_t)y * im.w + x);
}
__device__ __forceinline__ float4 load_clamp(const Img& im, int x, int y)
{
    x = min(max(x, 0), im.w - 1);
    y = min(max(y, 0), im.h - 1);
    return load_px(im.p, im.type, (size_t)y * im.w + x);
}

// Approximate SFU operations for the tolerance paths (1-2 ulp; inputs in the normal range)
__device__ __forceinline__ float rcp_approx(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float rsqrt_approx(float x)
{
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float ex2_approx(float x)
{
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

} // namespace tc
