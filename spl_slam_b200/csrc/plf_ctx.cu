// plf_ctx.cu -- context, stream, timer and scratch management for libplf.so.
#include "plf_common.cuh"

extern "C" plf_status plf_ctx_create(int device, plf_ctx** out)
{
    if (!out) return PLF_ERR_INVALID;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return PLF_ERR_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return PLF_ERR_CUDA;
    plf_ctx* c = (plf_ctx*)calloc(1, sizeof(plf_ctx));
    if (!c) return PLF_ERR_INVALID;
    c->device = device;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&c->ev0) != cudaSuccess || cudaEventCreate(&c->ev1) != cudaSuccess) {
        free(c);
        return PLF_ERR_CUDA;
    }
    *out = c;
    return PLF_OK;
}

extern "C" void plf_ctx_destroy(plf_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->scratch) cudaFree(c->scratch);
    if (c->pinned) cudaFreeHost(c->pinned);
    cudaEventDestroy(c->ev0);
    cudaEventDestroy(c->ev1);
    cudaStreamDestroy(c->stream);
    free(c);
}

extern "C" const char* plf_last_error(const plf_ctx* c) { return c ? c->err : "null context"; }

extern "C" plf_status plf_ctx_synchronize(plf_ctx* c)
{
    if (!c) return PLF_ERR_INVALID;
    PLF_CUDA(c, cudaStreamSynchronize(c->stream));
    return PLF_OK;
}

extern "C" void* plf_ctx_stream(plf_ctx* c) { return c ? (void*)c->stream : nullptr; }

extern "C" plf_status plf_timer_start(plf_ctx* c)
{
    if (!c) return PLF_ERR_INVALID;
    PLF_CUDA(c, cudaEventRecord(c->ev0, c->stream));
    return PLF_OK;
}

extern "C" plf_status plf_timer_stop(plf_ctx* c, float* ms)
{
    if (!c || !ms) return PLF_ERR_INVALID;
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
