// plf_shard.cu -- multi-GPU brute-force matching behind the C ABI: the train set is row-sharded over the GPUs of
// one node, one process per GPU; every rank computes its local top-2 with GLOBAL train indices, the per-shard tables are
// exchanged with ONE ncclAllGather (16 bytes per query per rank over NVLink / NVSwitch) on the context's stream, and a merge
// kernel takes the two smallest by (distance, index) -- the single-GPU cv::BFMatcher ordering, ties included.  Everything is
// queued on the context stream: no host synchronisation between the local search, the collective and the merge.
//
// NCCL is bound at run time (dlopen of libnccl.so.2, the copy the process already has -- torch's when the host is Python):
// libplf.so keeps no link-time dependency on it, and a single-GPU user never loads it.
#include "plf_common.cuh"
#ifndef PLF_EMU
#include <dlfcn.h>
#include <mutex>

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclInt32_ = 2 };   // ncclDataType_t: ncclInt8 0, ncclUint8 1, ncclInt32 2

namespace {
struct NcclApi {
    void* lib;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*);
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t);
    const char* (*GetErrorString)(ncclResult_t);
    const char* err;
} g_nccl;
std::once_flag g_nccl_once;

void nccl_bind()
{
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.lib) break;
    }
    if (!g_nccl.lib) { g_nccl.err = "libnccl.so.2 not found (dlopen)"; return; }
#define BIND(field, sym) \
    *(void**)&g_nccl.field = dlsym(g_nccl.lib, sym); \
    if (!g_nccl.field) { g_nccl.err = "symbol " sym " missing in libnccl"; return; }
    BIND(GetUniqueId, "ncclGetUniqueId")
    BIND(CommInitRank, "ncclCommInitRank")
    BIND(CommDestroy, "ncclCommDestroy")
    BIND(AllGather, "ncclAllGather")
    BIND(GetErrorString, "ncclGetErrorString")
#undef BIND
}
const char* nccl_ready()
{
    std::call_once(g_nccl_once, nccl_bind);
    return g_nccl.err;
}
}

struct plf_comm {
    plf_ctx* ctx;
    ncclComm_t comm;
    int rank, world;
    int32_t* d_parts;       // [world][nq][2] idx, then [world][nq][2] dist
    int32_t* d_local;       // [nq][2] idx, [nq][2] dist
    int cap_q;
};

#define PLF_NCCL(ctx, call)                                                                                          \
    do {                                                                                                             \
        ncclResult_t r_ = (call);                                                                                    \
        if (r_ != 0) return plf_fail((ctx), PLF_ERR_CUDA, "%s failed: %s", #call, g_nccl.GetErrorString(r_));        \
    } while (0)

extern "C" plf_status plf_comm_unique_id(uint8_t id[128])
{
    if (!id) return PLF_ERR_INVALID;
    if (nccl_ready()) return PLF_ERR_CUDA;
    ncclUniqueId u;
    if (g_nccl.GetUniqueId(&u) != 0) return PLF_ERR_CUDA;
    memcpy(id, u.internal, 128);
    return PLF_OK;
}

extern "C" plf_status plf_comm_create(plf_ctx* ctx, const uint8_t id[128], int rank, int world, plf_comm** out)
{
    if (!ctx || !id || !out || world < 1 || rank < 0 || rank >= world) return plf_fail(ctx, PLF_ERR_INVALID, "plf_comm_create: bad arguments");
    if (const char* e = nccl_ready()) return plf_fail(ctx, PLF_ERR_CUDA, "NCCL unavailable: %s", e);
    PLF_CUDA(ctx, cudaSetDevice(ctx->device));
    plf_comm* c = (plf_comm*)calloc(1, sizeof(plf_comm));
    if (!c) return plf_fail(ctx, PLF_ERR_CUDA, "out of memory");
    c->ctx = ctx; c->rank = rank; c->world = world;
    ncclUniqueId u;
    memcpy(u.internal, id, 128);
    ncclResult_t r = g_nccl.CommInitRank(&c->comm, world, u, rank);
    if (r != 0) { free(c); return plf_fail(ctx, PLF_ERR_CUDA, "ncclCommInitRank failed: %s", g_nccl.GetErrorString(r)); }
    *out = c;
    return PLF_OK;
}

extern "C" void plf_comm_destroy(plf_comm* c)
{
    if (!c) return;
    cudaSetDevice(c->ctx->device);
    if (c->comm) g_nccl.CommDestroy(c->comm);
    cudaFree(c->d_parts);
    cudaFree(c->d_local);
    free(c);
}

extern "C" int plf_comm_rank(const plf_comm* c) { return c ? c->rank : -1; }
extern "C" int plf_comm_world(const plf_comm* c) { return c ? c->world : 0; }

extern "C" plf_status plf_hamming_knn2_sharded_device(plf_ctx* ctx, plf_comm* c, const uint8_t* dev_q, int nq, const uint8_t* dev_t_local,
                                                      int64_t nt_local, int64_t train_index_base, int32_t* dev_idx, int32_t* dev_dist)
{
    if (!ctx || !c || c->ctx != ctx || nq < 0 || nt_local < 0 || (nq > 0 && (!dev_q || !dev_idx || !dev_dist)))
        return plf_fail(ctx, PLF_ERR_INVALID, "plf_hamming_knn2_sharded_device: bad arguments");
    if (nq == 0) return PLF_OK;
    PLF_CUDA(ctx, cudaSetDevice(ctx->device));
    if (c->world == 1) return plf_hamming_knn2_device(ctx, dev_q, nq, dev_t_local, nt_local, train_index_base, dev_idx, dev_dist);
    if (nq > c->cap_q) {
        cudaFree(c->d_parts); cudaFree(c->d_local);
        c->d_parts = c->d_local = nullptr; c->cap_q = 0;
        PLF_CUDA(ctx, cudaMalloc((void**)&c->d_parts, (size_t)c->world * nq * 4 * sizeof(int32_t)));
        PLF_CUDA(ctx, cudaMalloc((void**)&c->d_local, (size_t)nq * 4 * sizeof(int32_t)));
        c->cap_q = nq;
    }
    int32_t* lidx = c->d_local;
    int32_t* ldst = c->d_local + (size_t)nq * 2;
    int32_t* pidx = c->d_parts;
    int32_t* pdst = c->d_parts + (size_t)c->world * nq * 2;
    plf_status st = plf_hamming_knn2_device(ctx, dev_q, nq, dev_t_local, nt_local, train_index_base, lidx, ldst);
    if (st) return st;
    PLF_NCCL(ctx, g_nccl.AllGather(lidx, pidx, (size_t)nq * 2, ncclInt32_, c->comm, ctx->stream));
    PLF_NCCL(ctx, g_nccl.AllGather(ldst, pdst, (size_t)nq * 2, ncclInt32_, c->comm, ctx->stream));
    return plf_knn2_merge_device(ctx, pidx, pdst, c->world, nq, dev_idx, dev_dist);
}

// Linematcher::matchNNR over a sharded train set: sharded top-2, then the ratio test (src/Linematcher.cc:534-538)
extern "C" plf_status plf_match_nnr_sharded_device(plf_ctx* ctx, plf_comm* c, const uint8_t* dev_q, int nq, const uint8_t* dev_t_local,
                                                   int64_t nt_local, int64_t train_index_base, float nnr, int32_t* dev_idx, int32_t* dev_dist,
                                                   int32_t* dev_matches12, int32_t* dev_nmatches)
{
    plf_status st = plf_hamming_knn2_sharded_device(ctx, c, dev_q, nq, dev_t_local, nt_local, train_index_base, dev_idx, dev_dist);
    if (st) return st;
    return plf_nnr_from_knn2_device(ctx, dev_idx, dev_dist, nq, nnr, dev_matches12, dev_nmatches);
}

#else   // PLF_EMU: no NCCL in the CPU emulation; the N > 1 host logic is covered by the gloo test (tests/test_sharded_gloo.py)
struct plf_comm { int unused; };
extern "C" plf_status plf_comm_unique_id(uint8_t*) { return PLF_ERR_CUDA; }
extern "C" plf_status plf_comm_create(plf_ctx* ctx, const uint8_t*, int, int, plf_comm**) { return plf_fail(ctx, PLF_ERR_CUDA, "NCCL is not part of the emulated build"); }
extern "C" void plf_comm_destroy(plf_comm*) {}
extern "C" int plf_comm_rank(const plf_comm*) { return -1; }
extern "C" int plf_comm_world(const plf_comm*) { return 0; }
extern "C" plf_status plf_hamming_knn2_sharded_device(plf_ctx* ctx, plf_comm*, const uint8_t*, int, const uint8_t*, int64_t, int64_t, int32_t*, int32_t*)
{ return plf_fail(ctx, PLF_ERR_CUDA, "NCCL is not part of the emulated build"); }
extern "C" plf_status plf_match_nnr_sharded_device(plf_ctx* ctx, plf_comm*, const uint8_t*, int, const uint8_t*, int64_t, int64_t, float, int32_t*, int32_t*,
                                                   int32_t*, int32_t*)
{ return plf_fail(ctx, PLF_ERR_CUDA, "NCCL is not part of the emulated build"); }
#endif
