// tc_api.cu -- runtime part of the C ABI: errors, device/stream management, the
// stream-ordered frame pool and pinned staging.  See include/toucan_cuda.h.
#include <tuple>
#include <atomic>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

#include "tc_tma.cuh"
#include "tc_common.cuh"

namespace tc {

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches { 0 };

int fail(int status, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return status;
}

int check_cuda(cudaError_t e, const char* what)
{
    if (e == cudaSuccess) return TC_OK;
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver)
        return fail(TC_ERR_NO_DEVICE, "%s: %s (no CUDA device; there is no CPU fallback)", what, cudaGetErrorString(e));
    return fail(TC_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

void count_launch(unsigned n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int num_sms()
{
    // one GPU model per box: the first caller's device answers for all (initialisation is thread-safe)
    static const int sms = [] {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        return n;
    }();
    return sms;
}

int ensure_dynamic_smem(const void* kernel, size_t bytes)
{
    static std::mutex mutex;
    static std::map<std::pair<const void*, int>, size_t> done;
    int device = 0;
    TC_CUDA(cudaGetDevice(&device));
    std::lock_guard<std::mutex> lock(mutex);
    size_t& have = done[std::make_pair(kernel, device)];
    if (bytes > have) {
        TC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        TC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        have = bytes;
    }
    return TC_OK;
}

int validate(const tc_image* im, const char* name)
{
    if (!im) return fail(TC_ERR_INVALID, "%s: null image", name);
    if (!im->data) return fail(TC_ERR_INVALID, "%s: null data pointer", name);
    if (im->w <= 0 || im->h <= 0) return fail(TC_ERR_INVALID, "%s: empty image %dx%d", name, im->w, im->h);
    if (im->type < TC_U8 || im->type > TC_F32) return fail(TC_ERR_INVALID, "%s: unknown pixel type %d", name, im->type);
    if ((reinterpret_cast<uintptr_t>(im->data) & 15u) != 0) return fail(TC_ERR_INVALID, "%s: data not 16-byte aligned", name);
    return TC_OK;
}

} // namespace tc

using namespace tc;


// ---- tensor maps (tc_tma.cuh) ------------------------------------------------------------------------
namespace tc {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled()
{
    static const EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            p = nullptr;
        }
        return (EncodeTiledFn)p;
    }();
    return fn;
}
bool tma_available() { return encode_tiled() != nullptr; }

// a render reads the same stills frame after frame: descriptors are cached by every argument
int tensor_map_2d(const void* base, int elem_bytes, unsigned long long dim0, unsigned long long dim1, unsigned long long pitch_bytes,
                  unsigned box0, unsigned box1, CUtensorMap* out)
{
    typedef std::tuple<const void*, int, unsigned long long, unsigned long long, unsigned long long, unsigned, unsigned> Key;
    static std::mutex m;
    static std::map<Key, CUtensorMap> cache;
    const Key key(base, elem_bytes, dim0, dim1, pitch_bytes, box0, box1);
    {
        std::lock_guard<std::mutex> lock(m);
        const auto it = cache.find(key);
        if (it != cache.end()) {
            *out = it->second;
            return TC_OK;
        }
    }
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return fail(TC_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled is not available");
    if (((uintptr_t)base & 15) || (pitch_bytes & 15) || box0 > 256 || box1 > 256 || ((unsigned long long)box0 * elem_bytes & 15))
        return fail(TC_ERR_INVALID, "tensor map: base, pitch and box row must be 16-byte multiples, box <= 256 elements per side");
    const cuuint64_t dims[2] = { dim0, dim1 };
    const cuuint64_t strides[1] = { pitch_bytes };
    const cuuint32_t box[2] = { box0, box1 };
    const cuuint32_t estr[2] = { 1, 1 };
    // plain copies: the element type only sets the element size (and zero as the out-of-bounds fill)
    const CUtensorMapDataType dt = elem_bytes == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_UINT32;
    CUtensorMap tm;
    const CUresult r = enc(&tm, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(TC_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    std::lock_guard<std::mutex> lock(m);
    if (cache.size() > 4096) cache.clear();
    cache[key] = tm;
    *out = tm;
    return TC_OK;
}
} // namespace tc

extern "C" {

const char* tc_version(void) { return "toucan_b200 0.1 (sm_100a)"; }
const char* tc_last_error(void) { return g_err; }

int tc_device_count(int* count)
{
    if (!count) return fail(TC_ERR_INVALID, "tc_device_count: null");
    *count = 0;
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) {
        *count = 0;
        return check_cuda(e, "cudaGetDeviceCount");
    }
    return TC_OK;
}

int tc_set_device(int device)
{
    TC_CUDA(cudaSetDevice(device));
    // keep freed blocks cached in the stream-ordered pool: no cudaMalloc in the frame loop
    cudaMemPool_t pool;
    TC_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
    unsigned long long thr = ~0ull;
    TC_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
    return TC_OK;
}

int tc_stream_create(tc_stream* out)
{
    if (!out) return fail(TC_ERR_INVALID, "tc_stream_create: null");
    cudaStream_t s;
    TC_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    *out = s;
    return TC_OK;
}
int tc_stream_destroy(tc_stream s)
{
    TC_CUDA(cudaStreamDestroy((cudaStream_t)s));
    return TC_OK;
}
int tc_stream_sync(tc_stream s)
{
    TC_CUDA(cudaStreamSynchronize((cudaStream_t)s));
    return TC_OK;
}

int tc_event_create(tc_event* out)
{
    if (!out) return fail(TC_ERR_INVALID, "tc_event_create: null");
    cudaEvent_t e;
    TC_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    *out = e;
    return TC_OK;
}
int tc_event_create_timing(tc_event* out)
{
    if (!out) return fail(TC_ERR_INVALID, "tc_event_create_timing: null");
    cudaEvent_t e;
    TC_CUDA(cudaEventCreate(&e));
    *out = e;
    return TC_OK;
}
int tc_event_elapsed_ms(tc_event start, tc_event stop, float* ms)
{
    if (!start || !stop || !ms) return fail(TC_ERR_INVALID, "tc_event_elapsed_ms: null");
    TC_CUDA(cudaEventElapsedTime(ms, (cudaEvent_t)start, (cudaEvent_t)stop));
    return TC_OK;
}
int tc_event_destroy(tc_event e)
{
    if (!e) return TC_OK;
    TC_CUDA(cudaEventDestroy((cudaEvent_t)e));
    return TC_OK;
}
int tc_event_record(tc_event e, tc_stream s)
{
    TC_CUDA(cudaEventRecord((cudaEvent_t)e, (cudaStream_t)s));
    return TC_OK;
}
int tc_event_sync(tc_event e)
{
    TC_CUDA(cudaEventSynchronize((cudaEvent_t)e));
    return TC_OK;
}
int tc_stream_wait_event(tc_stream s, tc_event e)
{
    TC_CUDA(cudaStreamWaitEvent((cudaStream_t)s, (cudaEvent_t)e, 0));
    return TC_OK;
}

size_t tc_type_size(int type)
{
    static const size_t sz[4] = { 1, 2, 2, 4 };
    return (type >= 0 && type < 4) ? sz[type] : 0;
}
size_t tc_image_bytes(int w, int h, int type)
{
    if (w <= 0 || h <= 0) return 0;
    return (size_t)w * (size_t)h * 4 * tc_type_size(type);
}
int tc_type_merge(int a, int b)
{
    if (a == b) return a;
    if (tc_type_size(a) < tc_type_size(b)) {
        int t = a;
        a = b;
        b = t;
    }
    if (a == TC_F32) return TC_F32;
    if ((a == TC_U16 || a == TC_F16) && b == TC_U8) return a;
    return TC_F32;
}

// host time spent inside the allocator (diagnostics: tc_alloc_stats)
static std::atomic<unsigned long long> g_alloc_calls{ 0 }, g_alloc_ns{ 0 };
struct AllocTimer {
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    ~AllocTimer()
    {
        g_alloc_ns += (unsigned long long)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
        ++g_alloc_calls;
    }
};

// ---- frame-buffer cache -----------------------------------------------------------------
// A render allocates the same handful of buffer sizes for every frame.  The CUDA stream-ordered pool
// serves them from its free list most of the time, but with several frame sizes in flight it can fall into
// re-mapping physical memory on every request (measured: 0.27 ms of host time per frame for a 3-launch 4K
// frame, ten times the cost of building and launching the frame).  So freed buffers are kept here by exact
// size, per device: a free records an event on the freeing stream, the next allocation of that size makes
// its stream wait for that event and takes the buffer -- no call into the pool in steady state.  The cache
// is bounded (TOUCAN_CACHE_BYTES, default 64 GiB of the 180 GB); beyond the bound buffers go back to the pool.
namespace {
struct CachedBlock { void* p; cudaEvent_t ev; };
struct DeviceCache {
    std::mutex m;
    std::map<size_t, std::vector<CachedBlock>> free_;
    std::map<void*, size_t> live;     // buffers handed out by tc_malloc
    std::vector<cudaEvent_t> events;  // recycled events
    size_t cached = 0;
};
constexpr int kMaxDevices = 64;
DeviceCache g_cache[kMaxDevices];

size_t cache_limit()
{
    // TOUCAN_CACHE_BYTES, else a third of the device's memory (60 GB of a B200's 180): the bound scales with the part
    static const size_t limit = [] {
        const char* env = getenv("TOUCAN_CACHE_BYTES");
        if (env && *env) return (size_t)strtoull(env, nullptr, 10);
        size_t free_b = 0, total = 0;
        if (cudaMemGetInfo(&free_b, &total) != cudaSuccess || !total) {
            cudaGetLastError();
            return (size_t)64 << 30;
        }
        return total / 3;
    }();
    return limit;
}

DeviceCache* current_cache()
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return nullptr;
    return &g_cache[dev];
}

// the device a pointer was allocated on (a buffer may be released by a thread bound to another GPU)
int owner_device(const void* p)
{
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) == cudaSuccess && attr.type == cudaMemoryTypeDevice) return attr.device;
    cudaGetLastError();
    return -1;
}

// give every cached buffer back to the stream-ordered pool (caller holds no lock)
void cache_flush(DeviceCache& c)
{
    std::lock_guard<std::mutex> lock(c.m);
    for (auto& kv : c.free_)
        for (CachedBlock& b : kv.second) {
            cudaEventSynchronize(b.ev);
            cudaFreeAsync(b.p, (cudaStream_t)0);
            c.events.push_back(b.ev);
        }
    c.free_.clear();
    c.cached = 0;
}
} // namespace

int tc_malloc(void** out, size_t bytes, tc_stream s)
{
    if (!out) return fail(TC_ERR_INVALID, "tc_malloc: null");
    *out = nullptr;
    if (bytes == 0) return TC_OK;
    AllocTimer timer;
    DeviceCache* c = current_cache();
    if (c) {
        CachedBlock b { nullptr, nullptr };
        {
            std::lock_guard<std::mutex> lock(c->m);
            const auto it = c->free_.find(bytes);
            if (it != c->free_.end() && !it->second.empty()) {
                b = it->second.back();
                it->second.pop_back();
                c->cached -= bytes;
                c->live[b.p] = bytes;
            }
        }
        if (b.p) {
            // everything the previous owner enqueued before freeing the buffer finishes first
            const cudaError_t e = cudaStreamWaitEvent((cudaStream_t)s, b.ev, 0);
            {
                std::lock_guard<std::mutex> lock(c->m);
                c->events.push_back(b.ev);
            }
            TC_CUDA(e);
            *out = b.p;
            return TC_OK;
        }
    }
    cudaError_t e = cudaMallocAsync(out, bytes, (cudaStream_t)s);
    if (cudaErrorMemoryAllocation == e && c) {
        // out of memory while buffers of other sizes sit idle in the cache: give them back to the pool and retry once
        cudaGetLastError();
        cudaDeviceSynchronize();
        cache_flush(*c);
        cudaDeviceSynchronize();
        e = cudaMallocAsync(out, bytes, (cudaStream_t)s);
    }
    TC_CUDA(e);
    if (c) {
        std::lock_guard<std::mutex> lock(c->m);
        c->live[*out] = bytes;
    }
    return TC_OK;
}
int tc_free(void* p, tc_stream s)
{
    if (!p) return TC_OK;
    AllocTimer timer;
    DeviceCache* c = current_cache();
    if (c) {
        bool mine;
        {
            std::lock_guard<std::mutex> lock(c->m);
            mine = c->live.count(p) != 0;
        }
        if (!mine) {
            int cur = 0;
            cudaGetDevice(&cur);
            const int owner = owner_device(p);
            if (owner >= 0 && owner != cur) return fail(TC_ERR_INVALID, "tc_free: the buffer belongs to another device (use tc_free_sync)");
        }
    }
    size_t bytes = 0;
    cudaEvent_t ev = nullptr;
    if (c) {
        std::lock_guard<std::mutex> lock(c->m);
        const auto it = c->live.find(p);
        if (it != c->live.end()) {
            bytes = it->second;
            c->live.erase(it);
            if (c->cached + bytes <= cache_limit()) {
                if (!c->events.empty()) {
                    ev = c->events.back();
                    c->events.pop_back();
                }
            } else {
                bytes = 0; // over the bound: straight back to the pool
            }
        }
    }
    if (bytes) {
        if (!ev && cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) ev = nullptr;
        if (ev && cudaEventRecord(ev, (cudaStream_t)s) == cudaSuccess) {
            std::lock_guard<std::mutex> lock(c->m);
            c->free_[bytes].push_back(CachedBlock { p, ev });
            c->cached += bytes;
            return TC_OK;
        }
        if (ev) {
            std::lock_guard<std::mutex> lock(c->m);
            c->events.push_back(ev);
        }
    }
    TC_CUDA(cudaFreeAsync(p, (cudaStream_t)s));
    return TC_OK;
}
void tc_alloc_stats(unsigned long long* calls, unsigned long long* nanoseconds)
{
    if (calls) *calls = g_alloc_calls.load();
    if (nanoseconds) *nanoseconds = g_alloc_ns.load();
}
int tc_free_sync(void* p)
{
    if (!p) return TC_OK;
    // look the buffer up in the cache of the device that owns it, not the calling thread's current device
    int cur = 0;
    cudaGetDevice(&cur);
    const int owner = owner_device(p);
    if (owner >= 0 && owner != cur) cudaSetDevice(owner);
    const int dev = owner >= 0 ? owner : cur;
    if (dev >= 0 && dev < kMaxDevices) {
        DeviceCache& c = g_cache[dev];
        std::lock_guard<std::mutex> lock(c.m);
        c.live.erase(p);
    }
    // cudaFree waits for all work on the device: safe for a buffer several streams have read
    const cudaError_t e = cudaFree(p);
    if (owner >= 0 && owner != cur) cudaSetDevice(cur);
    TC_CUDA(e);
    return TC_OK;
}
int tc_pool_trim(void)
{
    int dev = 0;
    TC_CUDA(cudaGetDevice(&dev));
    cudaMemPool_t pool;
    TC_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
    TC_CUDA(cudaDeviceSynchronize());
    if (dev >= 0 && dev < kMaxDevices) cache_flush(g_cache[dev]);
    TC_CUDA(cudaDeviceSynchronize());
    TC_CUDA(cudaMemPoolTrimTo(pool, 0));
    return TC_OK;
}
int tc_host_register(void* p, size_t bytes)
{
    if (!p || !bytes) return fail(TC_ERR_INVALID, "tc_host_register: null");
    TC_CUDA(cudaHostRegister(p, bytes, cudaHostRegisterPortable));
    return TC_OK;
}
int tc_host_unregister(void* p)
{
    if (!p) return TC_OK;
    TC_CUDA(cudaHostUnregister(p));
    return TC_OK;
}
int tc_host_alloc(void** out, size_t bytes)
{
    if (!out) return fail(TC_ERR_INVALID, "tc_host_alloc: null");
    TC_CUDA(cudaHostAlloc(out, bytes, cudaHostAllocDefault));
    return TC_OK;
}
int tc_host_free(void* p)
{
    if (!p) return TC_OK;
    TC_CUDA(cudaFreeHost(p));
    return TC_OK;
}
int tc_upload(void* dst, const void* src, size_t bytes, tc_stream s)
{
    if (!dst || !src) return fail(TC_ERR_INVALID, "tc_upload: null pointer");
    TC_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)s));
    return TC_OK;
}
int tc_download(void* dst, const void* src, size_t bytes, tc_stream s)
{
    if (!dst || !src) return fail(TC_ERR_INVALID, "tc_download: null pointer");
    TC_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)s));
    return TC_OK;
}
int tc_memset(void* dst, int value, size_t bytes, tc_stream s)
{
    if (!dst) return fail(TC_ERR_INVALID, "tc_memset: null pointer");
    TC_CUDA(cudaMemsetAsync(dst, value, bytes, (cudaStream_t)s));
    return TC_OK;
}
unsigned long long tc_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

} // extern "C"
