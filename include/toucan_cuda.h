/*
 * toucan_cuda.h -- C ABI of libtoucan_cuda.so: the sm_100a kernels behind toucan's
 * per-frame render path.
 *
 * This is the drop-in boundary for the pixel work.  In the reference every effect
 * plug-in's _render() body is a call into OpenImageIO::ImageBufAlgo on host memory;
 * each entry point below replaces one of those call sites (cited per function, paths
 * relative to the reference tree) and runs on DEVICE memory, stream-ordered.
 *
 * Conventions
 *   - images are interleaved RGBA (4 channels), row-major, top row first, tightly
 *     packed (row bytes = w * 4 * sizeof(type)), as produced by
 *     lib/toucanRender/PropertySet.cpp:431-471 (bufToPropSet).
 *   - tc_image::data is a device pointer, 16-byte aligned.
 *   - arithmetic is float32; storage conversion happens at load/store with OIIO's rules
 *     (u8/u16: round-half-up with clamp; half: round-to-nearest-even).
 *   - reads outside a source image return 0 (OIIO black wrap) unless stated.
 *   - every call is asynchronous on `stream` (a cudaStream_t) and returns TC_OK or a
 *     negative tc_status; tc_last_error() gives the message for the calling thread.
 *   - no CPU fallback exists: without a CUDA device every compute call fails.
 */
#ifndef TOUCAN_CUDA_H
#define TOUCAN_CUDA_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum tc_type { TC_U8 = 0, TC_U16 = 1, TC_F16 = 2, TC_F32 = 3 } tc_type;

typedef enum tc_status {
    TC_OK = 0,
    TC_ERR_INVALID = -1,     /* bad argument (null image, size <= 0, unknown name, ...) */
    TC_ERR_CUDA = -2,        /* a CUDA runtime call failed; see tc_last_error() */
    TC_ERR_UNSUPPORTED = -3, /* valid in the reference but outside this build's scope */
    TC_ERR_NO_DEVICE = -4
} tc_status;

typedef struct tc_image {
    void* data; /* device pointer */
    int w, h;
    int type;   /* tc_type */
} tc_image;

typedef void* tc_stream; /* cudaStream_t */
typedef void* tc_event;  /* cudaEvent_t */

/* ---- runtime ---------------------------------------------------------------------- */
const char* tc_version(void);
const char* tc_last_error(void);
int tc_device_count(int* count);
int tc_set_device(int device);
int tc_stream_create(tc_stream* out);
int tc_stream_destroy(tc_stream s);
int tc_stream_sync(tc_stream s);
/* events: hand a finished frame from the render stream to the download stream / the writer
   (the frame scheduler that replaces the serial loop of bin/toucan-render/App.cpp:222-264) */
int tc_event_create(tc_event* out);
int tc_event_destroy(tc_event e);
int tc_event_record(tc_event e, tc_stream s);
int tc_event_sync(tc_event e);
int tc_stream_wait_event(tc_stream s, tc_event e);
/* timing events (diagnostics: the per-node timers of toucan-render -timers; no reference counterpart) */
int tc_event_create_timing(tc_event* out);
int tc_event_elapsed_ms(tc_event start, tc_event stop, float* ms);
size_t tc_type_size(int type);
size_t tc_image_bytes(int w, int h, int type);
/* result type of op(A,B) into an uninitialised destination (OIIO TypeDesc::basetype_merge) */
int tc_type_merge(int a, int b);

/* stream-ordered device pool (replaces the per-node malloc+memset of
   lib/toucanRender/ImageEffect.cpp:93,108,126).  Memory is NOT zeroed. */
int tc_malloc(void** out, size_t bytes, tc_stream s);
int tc_free(void* p, tc_stream s);
int tc_free_sync(void* p); /* cudaFree: for memory whose stream no longer exists */
int tc_pool_trim(void);
/* pinned host staging + copies (upload of decoded frames: Read.cpp:64-101; download
   for the writers: bin/toucan-render/App.cpp:238-262) */
/* page-lock / unlock memory the caller owns (decoded source frames), so uploads run at PCIe speed from every GPU */
int tc_host_register(void* p, size_t bytes);
int tc_host_unregister(void* p);
int tc_host_alloc(void** out, size_t bytes);
int tc_host_free(void* p);
int tc_upload(void* dst_dev, const void* src_host, size_t bytes, tc_stream s);
int tc_download(void* dst_host, const void* src_dev, size_t bytes, tc_stream s);
int tc_memset(void* dst_dev, int value, size_t bytes, tc_stream s);
/* count of kernels launched by this library in the calling process (bench evidence) */
unsigned long long tc_launch_count(void);
/* diagnostics: calls into tc_malloc / tc_free so far and the host time spent inside them */
void tc_alloc_stats(unsigned long long* calls, unsigned long long* nanoseconds);

/* ---- generators: plugins/GeneratorPlugin.cpp -------------------------------------- */
/* ImageBufAlgo::fill(out, color)                          GeneratorPlugin.cpp:296-303 */
int tc_fill(const tc_image* dst, const float color[4], tc_stream s);
/* ImageBufAlgo::checker(out, cw, ch, 1, c1, c2)           GeneratorPlugin.cpp:205-221 */
int tc_checkers(const tc_image* dst, int cw, int ch, const float c1[4], const float c2[4], tc_stream s);
/* fill(out, top, bottom) | u8 temp + rotate270 + copy     GeneratorPlugin.cpp:393-431 */
int tc_gradient(const tc_image* dst, const float c1[4], const float c2[4], int vertical, tc_stream s);
/* ImageBufAlgo::noise(out, type, a, b, mono, seed) on a zeroed buffer  GeneratorPlugin.cpp:525-531
   type: "gaussian"|"normal"|"white"|"uniform"|"salt" */
int tc_noise(const tc_image* dst, const char* type, float a, float b, int mono, int seed, tc_stream s);

/* ---- filters: plugins/FilterPlugin.cpp -------------------------------------------- */
/* make_kernel("gaussian", r, r) + convolve(out, src, k, true)   FilterPlugin.cpp:188-202 */
int tc_blur(const tc_image* dst, const tc_image* src, float radius, tc_stream s);
/* color_map(out, src, -1, name) + copy alpha              FilterPlugin.cpp:271-295 */
int tc_color_map(const tc_image* dst, const tc_image* src, const char* map_name, tc_stream s);
/* invert(out, src, ch 0..2) + copy alpha                  FilterPlugin.cpp:335-361 */
int tc_invert(const tc_image* dst, const tc_image* src, tc_stream s);
/* pow(out, src, value) on all 4 channels                  FilterPlugin.cpp:430-438 */
int tc_pow(const tc_image* dst, const tc_image* src, float value, tc_stream s);
/* saturate(out, src, value, 0)                            FilterPlugin.cpp:507-516 */
int tc_saturate(const tc_image* dst, const tc_image* src, float value, tc_stream s);
/* unsharp_mask(out, src, kernel, width, contrast, threshold)   FilterPlugin.cpp:607-618 */
int tc_unsharp_mask(const tc_image* dst, const tc_image* src, const char* kernel, float width, float contrast,
                    float threshold, tc_stream s);

/* ---- colour: plugins/ColorPlugin.cpp ---------------------------------------------- */
/* colorconvert(out, src, from, to, premult=0, key, val, cfg)   ColorPlugin.cpp:220-248.
   Baked decode-curve -> 3x3 -> encode-curve for the colour spaces of OCIO's built-in config. */
int tc_color_convert(const tc_image* dst, const tc_image* src, const char* fromspace, const char* tospace, tc_stream s);
/* same with OIIO's `unpremult` flag (toucan's "premult" parameter, ColorPlugin.cpp:241): pixels with
   alpha != 0 are divided by alpha before the transform and multiplied by it afterwards */
int tc_color_convert2(const tc_image* dst, const tc_image* src, const char* fromspace, const char* tospace, int unpremult, tc_stream s);
int tc_color_space_known(const char* name);
/* premult(out, src)                                       ColorPlugin.cpp:286 */
int tc_premult(const tc_image* dst, const tc_image* src, tc_stream s);
/* unpremult(out, src)                                     ColorPlugin.cpp:324 */
int tc_unpremult(const tc_image* dst, const tc_image* src, tc_stream s);

/* ---- transforms: plugins/TransformPlugin.cpp -------------------------------------- */
/* cut(src, ROI(pos, pos+size)) + copy(out, crop)          TransformPlugin.cpp:197-200 */
int tc_crop(const tc_image* dst, const tc_image* src, int px, int py, int cw, int ch, tc_stream s);
/* flip(out, src)                                          TransformPlugin.cpp:239-246 */
int tc_flip(const tc_image* dst, const tc_image* src, tc_stream s);
/* flop(out, src)                                          TransformPlugin.cpp:284-291 */
int tc_flop(const tc_image* dst, const tc_image* src, tc_stream s);
/* resize(out, src, filter, width, ROI(0,w,0,h))           TransformPlugin.cpp:374-379,
   lib/toucanRender/Comp.cpp:51-55, TransitionPlugin.cpp:226-230.
   filter_name: "" (default choice) | "lanczos3" | "blackman-harris" */
int tc_resize(const tc_image* dst, const tc_image* src, const char* filter_name, float filter_width, tc_stream s);
/* rotate(out, src, angle_radians, filter, width)          TransformPlugin.cpp:462-467 */
int tc_rotate(const tc_image* dst, const tc_image* src, float angle_radians, const char* filter_name,
              float filter_width, tc_stream s);

/* ---- transitions: plugins/TransitionPlugin.cpp ------------------------------------ */
/* fill, fill, mul, mul, add in ONE pass (sizes equal)     TransitionPlugin.cpp:242-264 */
int tc_dissolve(const tc_image* dst, const tc_image* from, const tc_image* to, double value, tc_stream s);
/* matte build + mul, mul, add in ONE pass                 TransitionPlugin.cpp:303-342 (vertical=0), :392-432 (1) */
int tc_wipe(const tc_image* dst, const tc_image* from, const tc_image* to, double value, int vertical, tc_stream s);

/* ---- compositing: lib/toucanRender/Comp.cpp --------------------------------------- */
/* premult (Comp.cpp:41, optional) + over (Comp.cpp:72) in ONE pass; fg, bg, dst same size */
int tc_over(const tc_image* dst, const tc_image* fg, const tc_image* bg, int premult_fg, tc_stream s);
/* the same over a CONSTANT background: the Fill generator at the bottom of a track stack (ImageGraph.cpp:165-171,
   GeneratorPlugin.cpp:283-306) as a colour instead of a frame.  `color` is the fill colour, `bg_type` the storage type
   the generator would have produced (its values are quantised to it); dst has fg's size. */
int tc_over_fill(const tc_image* dst, const tc_image* fg, const float color[4], int bg_type, int premult_fg, tc_stream s);
/* paste(dst, ox, oy, src) into a ZEROED dst of another size/type   Comp.cpp:56-67.
   Writes every dst pixel (0 outside the pasted rectangle). */
int tc_paste_on_black(const tc_image* dst, int ox, int oy, const tc_image* src, tc_stream s);
/* storage-type conversion / copy, 0 outside src           bin/toucan-render/App.cpp:285-291 */
int tc_convert(const tc_image* dst, const tc_image* src, tc_stream s);

/* raw output packing on the device, so the download moves the writer's bytes and not the
   working format's: paste(tmp(rawSpec), 0, 0, 0, 0, buf)    bin/toucan-render/App.cpp:267-300
   (rgb24, rgba, rgb48, rgba64, rgbaf16, rgbf32, rgbaf32 = `channels` x `type`, tightly packed).
   channels: 3 (alpha dropped) or 4. */
int tc_pack_raw(void* dst_dev, const tc_image* src, int type, int channels, tc_stream s);

/* ---- pointwise chain (cross-node fusion of consecutive pointwise effects) --------------
 * One pass for a chain of 2..TC_CHAIN_MAX pointwise graph nodes (Invert, Pow, Saturate, ColorMap, Premult,
 * Unpremult, ColorConvert; Flip / Flop, which only change where the source pixel is read and commute with every pointwise op;
 * CompNode's premult + over onto the Fill generator's colour, TC_CHAIN_OVER_FILL) applied bottom to top: ops[0] runs first.  Between two steps the value goes through
 * the storage type of the images (the reference materialises a typed buffer at every node boundary), so the
 * result is bit-identical to running the per-node kernels one after the other -- minus 2(n-1) frame passes.
 * name_a / name_b: ColorMap map name; ColorConvert from / to colour space.  flag: ColorConvert's unpremult.
 * An op the per-node call would refuse (unknown colour space) yields zeros, like its per-node kernel. */
enum { TC_CHAIN_INVERT = 0, TC_CHAIN_POW, TC_CHAIN_SATURATE, TC_CHAIN_COLOR_MAP, TC_CHAIN_PREMULT, TC_CHAIN_UNPREMULT, TC_CHAIN_COLOR_CONVERT,
       TC_CHAIN_ZERO,
       TC_CHAIN_OVER_FILL, /* lib/toucanRender/Comp.cpp:29-84 over the Fill generator's frame: color = the fill colour, flag = premult */
       TC_CHAIN_FLIP,      /* plugins/TransformPlugin.cpp:232-248 */
       TC_CHAIN_FLOP,      /* :277-293 */
       /* A chain may START with one convolution node (ops[0] only): the remaining ops run as the Blur / UnsharpMask kernel's
          epilogue when they are simple (Invert, Premult, Pow 2 / 3, over-Fill), else as a second launch over a temporary. */
       TC_CHAIN_BLUR,      /* plugins/FilterPlugin.cpp:178-205: value = radius */
       TC_CHAIN_UNSHARP }; /* :590-621: name_a = kernel, value = width, color[0] = contrast, color[1] = threshold */
#define TC_CHAIN_MAX 8
typedef struct tc_chain_op {
    int kind;
    float value;
    int flag;
    char name_a[96];
    char name_b[96];
    float color[4]; /* TC_CHAIN_OVER_FILL */
    int w, h;       /* TC_CHAIN_OVER_FILL: the Fill frame's size (the chain's frames must have it) */
} tc_chain_op;
/* what a plug-in appends to when the host asks it to describe its pixel operation instead of launching it
   (render-argument property "ToucanPropPointwiseChain") */
typedef struct tc_chain {
    int n;
    tc_chain_op ops[TC_CHAIN_MAX];
} tc_chain;
int tc_pointwise_chain(const tc_image* dst, const tc_image* src, const tc_chain_op* ops, int n, tc_stream s);

/* ---- draw (plugins/DrawPlugin.cpp) ---------------------------------------------------
   copy(src -> dst) + ImageBufAlgo::render_box / render_line: "colour over pixel"
   r[c] = color[c] + r[c] * (1 - color[3]) (plain store when color[3] == 1), clipped to the image.
   Box: corners inclusive; filled, or four skip-first lines; a single point when the corners coincide.
   Line: integer Bresenham walk from (x1, y1) to (x2, y2).  DrawPlugin.cpp:198-233, :306-341 */
int tc_draw_box(const tc_image* dst, const tc_image* src, int x1, int y1, int x2, int y2, const float color[4], int fill, tc_stream s);
int tc_draw_line(const tc_image* dst, const tc_image* src, int x1, int y1, int x2, int y2, const float color[4], int skip_first_point,
                 tc_stream s);

/* y4m output conversion on the device                     bin/toucan-render/App.cpp:37-43,340-485
   The reference pastes the frame into a u8/u16 RGB(A) buffer (ImageBufAlgo::paste) and runs
   libswscale (FFmpeg 8.0, SWS_FAST_BILINEAR) RGB24->YUV422P | RGB24->YUV444P | RGBA->YUVA444P |
   RGB48->YUV444P16, then writes the planes back to back after "FRAME\n".  tc_rgb_to_y4m writes
   exactly those bytes (Y, U, V[, A] planes, tightly packed, little-endian for 444p16) into
   dst_dev, which holds tc_y4m_frame_bytes(format, w, h) bytes.  Bit-exact with libswscale
   9.1.100 on x86-64.  Even widths >= 8 only. */
enum { TC_Y4M_422 = 0, TC_Y4M_444 = 1, TC_Y4M_444ALPHA = 2, TC_Y4M_444P16 = 3 };
size_t tc_y4m_frame_bytes(int format, int w, int h);
int tc_rgb_to_y4m(void* dst_dev, const tc_image* src, int format, tc_stream s);

/* ---- fused track stack (cross-node fusion of the hot chain) ------------------------
 * One pass over the output for a whole stack of tracks:
 *   acc = background colour
 *   for each layer (bottom to top):
 *       a = blur_a ? blur(src_a, radius_a) : src_a
 *       if (src_b) { b = blur_b ? blur(src_b, radius_b) : src_b;  a = dissolve(a, b, value) }
 *       acc = over(premult(a), acc)
 * i.e. ImageGraph.cpp:165-198 (_track + CompNode chain) with toucan:Blur
 * (FilterPlugin.cpp:188-202) and toucan:Dissolve (TransitionPlugin.cpp:242-264) per track.
 * All sources float32 and of the output size. */
typedef struct tc_stack_layer {
    const void* src_a; /* device, float32 RGBA, w*h */
    const void* src_b; /* optional second source (transition), or NULL */
    float radius_a;    /* toucan:Blur radius on a (0 = none) */
    float radius_b;
    double value;      /* dissolve value */
} tc_stack_layer;
int tc_stack_f32(const tc_image* dst, const float background[4], const tc_stack_layer* layers, int n_layers,
                 tc_stream s);

#ifdef __cplusplus
}
#endif
#endif /* TOUCAN_CUDA_H */
