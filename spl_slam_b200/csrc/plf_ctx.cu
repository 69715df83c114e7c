// plf_ctx.cu -- context, stream, timer and scratch management for libplf.so.
#include "plf_common.cuh"
#include <unistd.h>

// priority: 0 = default, < 0 = lower than default (filler work), > 0 = higher (latency-critical work); mapped onto the
// device's stream priority range
extern "C" plf_status plf_ctx_create_prio(int device, int priority, plf_ctx** out)
{
    if (!out) return PLF_ERR_INVALID;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return PLF_ERR_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return PLF_ERR_CUDA;
    plf_ctx* c = (plf_ctx*)calloc(1, sizeof(plf_ctx));
    if (!c) return PLF_ERR_INVALID;
    c->device = device;
    int lo = 0, hi = 0;   // numerically: lo = least priority (largest value), hi = greatest priority (smallest value)
#ifndef PLF_EMU
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
#endif
    const int prio = priority > 0 ? hi : (priority < 0 ? lo : (lo + hi) / 2);
    cudaError_t e;
#ifndef PLF_EMU
    e = cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prio);
#else
    (void)prio;
    e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
#endif
    if (e != cudaSuccess || cudaEventCreate(&c->ev0) != cudaSuccess || cudaEventCreate(&c->ev1) != cudaSuccess) {
        free(c);
        return PLF_ERR_CUDA;
    }
    *out = c;
    return PLF_OK;
}

extern "C" plf_status plf_ctx_create(int device, plf_ctx** out) { return plf_ctx_create_prio(device, 0, out); }

static void plf_prof_free(plf_ctx* c);

extern "C" void plf_ctx_destroy(plf_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    plf_prof_free(c);
    if (c->scratch) cudaFree(c->scratch);
    if (c->ioscratch) cudaFree(c->ioscratch);
    if (c->pinned) cudaFreeHost(c->pinned);
    cudaEventDestroy(c->ev0);
    cudaEventDestroy(c->ev1);
    cudaStreamDestroy(c->stream);
    free(c);
}

extern "C" const char* plf_last_error(const plf_ctx* c) { return c ? c->err : "null context"; }

plf_status plf_sync(plf_ctx* c, cudaStream_t st)
{
#ifndef PLF_EMU
    static const bool never_sleep = getenv("PLF_SPIN_SYNC") != nullptr;
    if (c->blocking && !never_sleep) {
        // sleep-and-poll: the waits of a batch call last milliseconds, 100 us of extra wake-up latency is noise, and the host
        // core is free for the other ranks' threads meanwhile
        for (;;) {
            const cudaError_t q = cudaStreamQuery(st);
            if (q == cudaSuccess) return PLF_OK;
            if (q != cudaErrorNotReady) return plf_fail(c, PLF_ERR_CUDA, "cudaStreamQuery failed: %s", cudaGetErrorString(q));
            (void)cudaGetLastError();       // "not ready" must not be mistaken for a launch error by the next PLF_CHECK_LAUNCH
            usleep(100);
        }
    }
#endif
    PLF_CUDA(c, cudaStreamSynchronize(st));
    return PLF_OK;
}

extern "C" plf_status plf_ctx_synchronize(plf_ctx* c)
{
    if (!c) return PLF_ERR_INVALID;
    PLF_CUDA(c, cudaSetDevice(c->device));
    return plf_sync(c, c->stream);
}

extern "C" void* plf_ctx_stream(plf_ctx* c) { return c ? (void*)c->stream : nullptr; }

extern "C" plf_status plf_timer_start(plf_ctx* c)
{
    if (!c) return PLF_ERR_INVALID;
    PLF_CUDA(c, cudaSetDevice(c->device));
    PLF_CUDA(c, cudaEventRecord(c->ev0, c->stream));
    return PLF_OK;
}

extern "C" plf_status plf_timer_stop(plf_ctx* c, float* ms)
{
    if (!c || !ms) return PLF_ERR_INVALID;
    PLF_CUDA(c, cudaSetDevice(c->device));
    PLF_CUDA(c, cudaEventRecord(c->ev1, c->stream));
    PLF_CUDA(c, cudaEventSynchronize(c->ev1));
    PLF_CUDA(c, cudaEventElapsedTime(ms, c->ev0, c->ev1));
    return PLF_OK;
}

extern "C" uint64_t plf_ctx_launch_count(const plf_ctx* c) { return c ? c->launches : 0; }

plf_status plf_ctx_scratch(plf_ctx* c, size_t bytes, void** out)
{
    if (bytes > c->scratch_bytes) {
        PLF_CUDA(c, cudaStreamSynchronize(c->stream));
        if (c->scratch) cudaFree(c->scratch);
        c->scratch = nullptr;
        c->scratch_bytes = 0;
        size_t want = plf_align_up(bytes + bytes / 4, 1 << 20);
        PLF_CUDA(c, cudaMalloc(&c->scratch, want));
        c->scratch_bytes = want;
    }
    *out = c->scratch;
    return PLF_OK;
}

plf_status plf_ctx_ioscratch(plf_ctx* c, size_t bytes, void** out)
{
    if (bytes > c->ioscratch_bytes) {
        PLF_CUDA(c, cudaStreamSynchronize(c->stream));
        if (c->ioscratch) cudaFree(c->ioscratch);
        c->ioscratch = nullptr;
        c->ioscratch_bytes = 0;
        size_t want = plf_align_up(bytes + bytes / 4, 1 << 20);
        PLF_CUDA(c, cudaMalloc(&c->ioscratch, want));
        c->ioscratch_bytes = want;
    }
    *out = c->ioscratch;
    return PLF_OK;
}

plf_status plf_ctx_pinned(plf_ctx* c, size_t bytes, void** out)
{
    if (bytes > c->pinned_bytes) {
        PLF_CUDA(c, cudaStreamSynchronize(c->stream));
        if (c->pinned) cudaFreeHost(c->pinned);
        c->pinned = nullptr;
        c->pinned_bytes = 0;
        size_t want = plf_align_up(bytes + bytes / 4, 1 << 20);
        PLF_CUDA(c, cudaMallocHost(&c->pinned, want));
        c->pinned_bytes = want;
    }
    *out = c->pinned;
    return PLF_OK;
}

// ---------------- per-kernel profiling (bench.py roofline: CUDA events on the launching stream) ----------------
#define PLF_PROF_MAX 8192
#define PLF_PROF_NAMES 64
struct plf_prof_state {
    cudaEvent_t ev[PLF_PROF_MAX][2];
    const char* name[PLF_PROF_MAX];
    int created;
    const char* names[PLF_PROF_NAMES];
    double total_ms[PLF_PROF_NAMES];
    long count[PLF_PROF_NAMES];
    int nnames;
};

// the profiling state and the events it created lazily
static void plf_prof_free(plf_ctx* c)
{
    if (!c->prof) return;
#ifndef PLF_EMU
    for (int i = 0; i < c->prof->created; i++) { cudaEventDestroy(c->prof->ev[i][0]); cudaEventDestroy(c->prof->ev[i][1]); }
#endif
    free(c->prof);
    c->prof = nullptr;
}

#ifndef PLF_EMU
void plf_prof_begin(plf_ctx* c, const char* name)
{
    if (!c->prof_on || c->prof_n >= PLF_PROF_MAX) return;
    plf_prof_state* p = c->prof;
    if (c->prof_n >= p->created) {
        cudaSetDevice(c->device);         // events are created on the current device
        cudaEventCreate(&p->ev[p->created][0]);
        cudaEventCreate(&p->ev[p->created][1]);
        p->created++;
    }
    p->name[c->prof_n] = name;
    cudaEventRecord(p->ev[c->prof_n][0], c->stream);
}
void plf_prof_end(plf_ctx* c)
{
    if (!c->prof_on || c->prof_n >= PLF_PROF_MAX) return;
    cudaEventRecord(c->prof->ev[c->prof_n][1], c->stream);
    c->prof_n++;
}
#endif

static void prof_collect(plf_ctx* c)
{
#ifndef PLF_EMU
    plf_prof_state* p = c->prof;
    if (!p) return;
    cudaStreamSynchronize(c->stream);
    for (int i = 0; i < c->prof_n; i++) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, p->ev[i][0], p->ev[i][1]) != cudaSuccess) continue;
        int k = 0;
        for (; k < p->nnames; k++) if (p->names[k] == p->name[i] || !strcmp(p->names[k], p->name[i])) break;
        if (k == p->nnames) {
            if (p->nnames >= PLF_PROF_NAMES) continue;
            p->names[k] = p->name[i]; p->total_ms[k] = 0; p->count[k] = 0; p->nnames++;
        }
        p->total_ms[k] += ms; p->count[k]++;
    }
    c->prof_n = 0;
#endif
}

// enable/disable per-launch event timing; enabling resets the accumulated table
extern "C" plf_status plf_profile_enable(plf_ctx* c, int on)
{
    if (!c) return PLF_ERR_INVALID;
    if (on) {
        if (!c->prof) c->prof = (plf_prof_state*)calloc(1, sizeof(plf_prof_state));
        c->prof->nnames = 0;
        c->prof_n = 0;
    } else if (c->prof_on) {
        prof_collect(c);
    }
    c->prof_on = on ? 1 : 0;
    return PLF_OK;
}

// writes "name total_ms launches\n" lines into buf; returns PLF_ERR_CAPACITY if buf is too small
extern "C" plf_status plf_profile_report(plf_ctx* c, char* buf, size_t bufsize)
{
    if (!c || !buf || bufsize < 2) return PLF_ERR_INVALID;
    buf[0] = 0;
    if (!c->prof) return PLF_OK;
    prof_collect(c);
    size_t o = 0;
    for (int k = 0; k < c->prof->nnames; k++) {
        int n = snprintf(buf + o, bufsize - o, "%s %.6f %ld\n", c->prof->names[k], c->prof->total_ms[k], c->prof->count[k]);
        if (n < 0 || (size_t)n >= bufsize - o) return PLF_ERR_CAPACITY;
        o += n;
    }
    return PLF_OK;
}

// diagnostic: "name start_ms end_ms" per recorded launch, times relative to `ref`'s last plf_timer_start event
// (consumes the pending records; call before plf_profile_report)
extern "C" plf_status plf_profile_timeline(plf_ctx* c, plf_ctx* ref, char* buf, size_t bufsize)
{
    if (!c || !ref || !buf || bufsize < 2) return PLF_ERR_INVALID;
    buf[0] = 0;
#ifndef PLF_EMU
    plf_prof_state* p = c->prof;
    if (!p) return PLF_OK;
    cudaStreamSynchronize(c->stream);
    size_t o = 0;
    for (int i = 0; i < c->prof_n; i++) {
        float t0 = 0, t1 = 0;
        if (cudaEventElapsedTime(&t0, ref->ev0, p->ev[i][0]) != cudaSuccess) continue;
        if (cudaEventElapsedTime(&t1, ref->ev0, p->ev[i][1]) != cudaSuccess) continue;
        int n = snprintf(buf + o, bufsize - o, "%s %.4f %.4f\n", p->name[i], t0, t1);
        if (n < 0 || (size_t)n >= bufsize - o) return PLF_ERR_CAPACITY;
        o += n;
    }
#endif
    return PLF_OK;
}

// make `ctx` wait (on the device) for everything queued so far on `other` (ORB and line contexts run on two streams)
extern "C" plf_status plf_ctx_wait(plf_ctx* c, plf_ctx* other)
{
    if (!c || !other) return PLF_ERR_INVALID;
    // the event must belong to the streams' device: a host thread that has never touched CUDA has device 0 current, which is the
    // wrong one on every rank but the first (an event of another device fails with "invalid resource handle")
    PLF_CUDA(c, cudaSetDevice(c->device));
    cudaEvent_t e;
    PLF_CUDA(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    cudaError_t r = cudaEventRecord(e, other->stream);
    if (r == cudaSuccess) r = cudaStreamWaitEvent(c->stream, e, 0);
    cudaEventDestroy(e);      // released once the wait has been consumed; also on the error path
    PLF_CUDA(c, r);
    return PLF_OK;
}

// One upload for several consumers (ORB and line extractor of a frame, src/Frame.cc:301-304): the images go to the device once,
// on this context's stream, in 16 MB pieces so that transfers queued by other contexts can slip in between; the consumers
// make their streams wait for it with plf_ctx_wait(consumer, this) and call the *_from_device entry points.
extern "C" plf_status plf_upload(plf_ctx* ctx, void* dev_dst, const void* host_src, size_t bytes)
{
    if (!ctx || !dev_dst || !host_src) return plf_fail(ctx, PLF_ERR_INVALID, "plf_upload: bad arguments");
    PLF_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t piece = (size_t)16 << 20;
    for (size_t off = 0; off < bytes; off += piece)
        PLF_CUDA(ctx, cudaMemcpyAsync((char*)dev_dst + off, (const char*)host_src + off, bytes - off < piece ? bytes - off : piece,
                                      cudaMemcpyHostToDevice, ctx->stream));
    return PLF_OK;
}
extern "C" plf_status plf_device_malloc(plf_ctx* ctx, size_t bytes, void** out)
{
    if (!ctx || !out) return PLF_ERR_INVALID;
    PLF_CUDA(ctx, cudaSetDevice(ctx->device));
    PLF_CUDA(ctx, cudaMalloc(out, bytes ? bytes : 1));
    return PLF_OK;
}
extern "C" void plf_device_free(plf_ctx* ctx, void* p)
{
    if (!ctx || !p) return;
    cudaSetDevice(ctx->device);
    cudaFree(p);
}
