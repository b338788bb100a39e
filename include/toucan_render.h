/*
 * toucan_render.h -- C ABI of libtoucan_render.so: the C++ host of the render path
 * (toucan::ImageEffectHost + toucan::TimelineWrapper + toucan::ImageGraph, the same classes
 * as the reference's lib/toucanRender) for callers that are not C++ -- the Python test and
 * bench harness here, a cgo/JNI/ctypes binding elsewhere.
 *
 * One call = one step of bin/toucan-render/App.cpp:152-265:
 *   tr_host_create        ImageEffectHost(context, searchPath)               App.cpp:199
 *   tr_timeline_open*     TimelineWrapper(path) + ImageGraph(context, ...)    App.cpp:162-170
 *   tr_timeline_render    graph->exec(host, time) ; node->exec() ; download   App.cpp:232-262
 * Every function returns 0 on success or a negative value; tr_last_error() holds the
 * message for the calling thread.  Rendering needs a CUDA device; there is no CPU path.
 */
#ifndef TOUCAN_RENDER_H
#define TOUCAN_RENDER_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tr_host tr_host;
typedef struct tr_timeline tr_timeline;

const char* tr_last_error(void);

/* Bind the calling thread to a CUDA device (one render thread per GPU). */
int tr_bind_device(int device);
/* Same, rendering on a stream the caller owns (a cudaStream_t). */
int tr_bind_stream(int device, void* stream);
/* Wait for everything the calling thread enqueued. */
int tr_sync(void);

/* OpenFX host: loads every *.ofx module under plugin_dir (NULL: the directory next to this
   library, then $OFX_PLUGIN_PATH).  Works without a GPU (describe only). */
int tr_host_create(const char* plugin_dir, tr_host** out);
void tr_host_destroy(tr_host* host);
int tr_host_plugin_count(const tr_host* host);
const char* tr_host_plugin_identifier(const tr_host* host, int index);

/* Timelines.  Each timeline owns a media store: decoded frames registered by resolved path. */
int tr_timeline_open(const char* otio_path, tr_timeline** out);
int tr_timeline_open_json(const char* json, const char* directory, tr_timeline** out);
void tr_timeline_close(tr_timeline* timeline);
/* Resolve a media URL the way TimelineWrapper::getMediaPath does. */
int tr_timeline_media_path(const tr_timeline* timeline, const char* url, char* out, size_t capacity);
/* Register a decoded RGBA frame (tightly packed, `type` = tc_type) under a resolved path.
   The pixels are BORROWED: host memory is uploaded on every render that reads it, device
   memory (device != 0) is used in place.  Must be called before tr_timeline_prepare. */
int tr_timeline_set_media(tr_timeline* timeline, const char* path, const void* pixels, int device, int width, int height, int type);
/* Device source cache: keep up to `bytes` of uploaded host frames resident per GPU (LRU), so a still
   is uploaded once instead of on every frame.  0 (default) = upload on every frame. */
int tr_timeline_set_source_cache(tr_timeline* timeline, size_t bytes);
/* Host -> device source uploads so far (count, bytes): what the device source cache saved shows up here. */
int tr_timeline_upload_stats(const tr_timeline* timeline, unsigned long long* uploads, unsigned long long* bytes);
/* Empty the device source cache (the next frame uploads what it reads again). */
int tr_timeline_drop_source_cache(tr_timeline* timeline);
/* Build the ImageGraph (reads the first clip's media information). */
int tr_timeline_prepare(tr_timeline* timeline);
/* start/rate of the first frame, number of frames (App.cpp:222-224), image size and tc_type. */
int tr_timeline_info(const tr_timeline* timeline, double* start_value, double* rate, int* frames, int* width, int* height, int* type);

/* The node tree ImageGraph::exec builds for a time, one node per line, inputs indented. */
int tr_timeline_describe(tr_timeline* timeline, tr_host* host, double time_value, char* out, size_t capacity);
/* Render a frame into host memory (blocks until the frame is complete).  width/height/type
   return the frame's format; capacity must hold it. */
int tr_timeline_render(tr_timeline* timeline, tr_host* host, double time_value, void* out, size_t capacity, int* width, int* height,
                       int* type);
/* Render a frame and leave it in device memory: *out is the device pointer of the finished
   frame, enqueued on the thread's stream (tr_sync before reading from another stream or
   the host).  The frame stays valid until the next render_resident call on this timeline. */
int tr_timeline_render_resident(tr_timeline* timeline, tr_host* host, double time_value, void** out, int* width, int* height, int* type);

/* The frame scheduler (replaces the serial loop of App.cpp:222-264): renders frames
   [first, last) of the timeline on the given GPUs -- contiguous blocks of frames per GPU, one
   worker thread + OpenFX host + ImageGraph per GPU, finished frames downloaded into pinned ring
   buffers while the next frame renders -- and calls `callback` on the calling thread once per
   frame in timeline order.  `pixels` is pinned host memory, valid during the callback. */
typedef void (*tr_frame_callback)(void* user, int frame, const void* pixels, int width, int height, int type);
int tr_render_range(tr_timeline* timeline, const char* plugin_dir, const int* devices, int n_devices, int first, int last,
                    int frames_in_flight, tr_frame_callback callback, void* user);

/* Same with the output stage of `toucan-render - -raw <format>` (App.cpp:267-300) on the device:
   raw_format = "rgb24" | "rgba" | "rgb48" | "rgba64" | "rgbaf16" | "rgbf32" | "rgbaf32" (NULL = the
   working format).  Frames whose storage type differs from the format's are converted before the
   download; the callback's `type` then carries the raw type, plus 0x100 for 3-channel formats. */
int tr_render_range_raw(tr_timeline* timeline, const char* plugin_dir, const int* devices, int n_devices, int first, int last,
                        int frames_in_flight, const char* raw_format, tr_frame_callback callback, void* user);

/* Same with the output stage of `toucan-render - -y4m <format>` (App.cpp:340-485) on the device:
   y4m_format = "422" | "444" | "444alpha" | "444p16".  Every frame is converted to the planar YUV bytes the
   reference writes after "FRAME\n" (Y, U, V[, A] planes back to back; tc_y4m_frame_bytes() of them) and only
   those are downloaded; the callback's `type` is 0x200 | TC_Y4M_*. */
int tr_render_range_y4m(tr_timeline* timeline, const char* plugin_dir, const int* devices, int n_devices, int first, int last,
                        int frames_in_flight, const char* y4m_format, tr_frame_callback callback, void* user);
/* The stream header of App.cpp:302-338, NUL-terminated, into out[capacity]. */
int tr_y4m_header(int width, int height, double rate, const char* y4m_format, char* out, size_t capacity);

/* The scheduler's partition: frames [first, last) -> the contiguous block [*lo, *hi) owned by
   worker `index` of `count` (ceil(N / G) frames per GPU). */
void tr_frame_block(int first, int last, int index, int count, int* lo, int* hi);

/* toucan::fit (lib/toucanRender/Util.cpp:123-147): the box, inside a frame of aw x ah, that a bw x bh image is fitted
   into (aspect preserved, centred).  out = { min x, min y, width, height }.  CompNode and the Dissolve / Wipe effects
   place a differently sized input with it. */
void tr_fit(int aw, int ah, int bw, int bh, int out[4]);

/* The built-in PNG reader used when no decoded frame was registered for a path (RGBA out,
   A = 1 appended to RGB, alpha associated on read like OIIO's reader).  out may be NULL to
   query the size. */
int tr_decode_png(const char* path, void* out, size_t capacity, int* width, int* height, int* type);
/* Reader option = OpenImageIO's "oiio:UnassociatedAlpha" open hint: keep != 0 leaves the unassociated alpha of PNG / TIFF
   files as stored; 0 (the default, what the reference gets by opening its media with no hints, Read.cpp:39) multiplies colour
   by alpha on read.  The reference's own film strips of the transition timelines were rendered with unassociated frames
   (INTEGRATION.md "PNG alpha").  Process-wide; also TOUCAN_UNASSOCIATED_ALPHA=1 and `toucan-render -unassociated_alpha`. */
void tr_set_unassociated_alpha(int keep);
int tr_get_unassociated_alpha(void);
/* The built-in JPEG reader (baseline / extended / progressive Huffman, 8-bit YCbCr or RGB, 1x-2x sampling):
   libjpeg's defaults -- islow IDCT, fancy upsampling, integer colour tables -- bit for bit, RGBA out with
   A = 255 (Read.cpp:86-94).  Same calling convention as tr_decode_png. */
int tr_decode_jpeg(const char* path, void* out, size_t capacity, int* width, int* height, int* type);
/* The built-in OpenEXR reader (scanline, single-part, R/G/B[/A] HALF or FLOAT, compression NONE / RLE / ZIPS / ZIP / PIZ):
   RGBA out in the widest channel type (*type = TC_F16 or TC_F32), A = 1 when the file has no alpha.  Same calling
   convention as tr_decode_png. */
int tr_decode_exr(const char* path, void* out, size_t capacity, int* width, int* height, int* type);
/* The built-in TIFF reader (classic, stripped, chunky RGB / RGBA; 8/16-bit unsigned or 32-bit float; uncompressed, LZW,
   deflate, PackBits; predictor 1 or 2): RGBA out in the file's sample type, A = 1 for RGB, unassociated alpha associated
   on read.  Same calling convention as tr_decode_png. */
int tr_decode_tiff(const char* path, void* out, size_t capacity, int* width, int* height, int* type);
/* The writer behind `toucan-render out.####.exr`: an RGBA half (TC_F16) or float (TC_F32) host frame as a scanline
   OpenEXR file, ZIP-compressed (zip != 0, what the command line tool writes) or uncompressed. */
int tr_write_exr(const char* path, const void* pixels, int width, int height, int type, int zip);

/* Cross-node fusion of the track stack (default on). */
void tr_set_fusion(int enabled);

/* Frame -> worker assignment of tr_render_range*: round-robin (default; all GPUs busy behind the one ordered writer) or
   contiguous blocks (tr_frame_block: consecutive frames per worker). */
void tr_set_interleave(int enabled);

/* Per-node device timers (the reference has none; `toucan-render -timers`): when on, every effect / composite / fused
   node brackets its kernels with two timing events.  tr_node_timers_report waits for the pending events, writes one line per
   node label ("label  calls  total ms  ms/call", largest first) and resets the totals; returns the bytes the text needs. */
void tr_set_node_timers(int enabled);
size_t tr_node_timers_report(char* out, size_t capacity);

#ifdef __cplusplus
}
#endif
#endif /* TOUCAN_RENDER_H */
