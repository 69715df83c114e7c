// plf_bow.cu -- DBoW2 vocabulary-tree descent on the device: the per-feature part of
// TemplatedVocabulary::transform(features, BowVector&, FeatureVector&, levelsup)
// (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1124-1190, single feature :1218-1258) as called from
// Frame::ComputeBoW (src/Frame.cc:724-731) with FORB::distance (Thirdparty/DBoW2/DBoW2/FORB.cpp:81-101).
// One warp per feature: at every level the (<= 32) children are scored by one lane each, the first minimum wins
// (the reference's strict `d < best_d` scan).  The BowVector / FeatureVector maps are a few hundred ordered
// insertions per frame and stay on the host (the shim replays addWeight / addFeature / normalize on the results).
#include "plf_common.cuh"
#include <vector>
#include <string>
#include <fstream>
#include <sstream>

struct plf_vocab {
    plf_ctx* ctx;
    int k, L, nnodes, nwords;
    int scoring, weighting;
    // device copies
    int* d_child_off;      // nnodes + 1
    int* d_child_ids;      // nnodes - 1 (every node but the root is somebody's child)
    uint8_t* d_desc;       // nnodes x 32
    int* d_word;           // word id per node (-1 for inner nodes)
    double* d_weight;      // per node
    // host copies kept for the map helpers
    std::vector<int> parent;
};

__global__ void __launch_bounds__(256)
k_bow_transform(const uint4* __restrict__ feat, int n, const int* __restrict__ child_off, const int* __restrict__ child_ids,
                const uint4* __restrict__ ndesc, const int* __restrict__ word, const double* __restrict__ weight, int nid_level,
                int* __restrict__ out_word, double* __restrict__ out_weight, int* __restrict__ out_node)
{
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= n) return;
    const uint4 a0 = feat[2 * (size_t)i], a1 = feat[2 * (size_t)i + 1];
    int final_id = 0, level = 0, nid = 0;
    for (;;) {
        const int c0 = child_off[final_id], nc = child_off[final_id + 1] - c0;
        if (nc <= 0) break;                       // leaf (isLeaf() == children.empty())
        ++level;
        int key = 0x7fffffff;
        if (lane < nc) {
            const int id = child_ids[c0 + lane];
            const uint4 b0 = ndesc[2 * (size_t)id], b1 = ndesc[2 * (size_t)id + 1];
            const int d = __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
                          __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
            key = (d << 5) | lane;                // first minimum in child order
        }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) key = min(key, __shfl_xor_sync(0xffffffffu, key, s));
        final_id = child_ids[c0 + (key & 31)];
        if (level == nid_level) nid = final_id;
    }
    if (lane == 0) {
        out_word[i] = word[final_id];
        out_weight[i] = weight[final_id];
        out_node[i] = nid;
    }
}

extern "C" plf_status plf_vocab_create(plf_ctx* ctx, int k, int L, int scoring, int weighting, int nnodes, const int32_t* parent,
                                       const uint8_t* desc, const double* weight, const uint8_t* is_leaf, plf_vocab** out)
{
    if (!ctx || !out) return PLF_ERR_INVALID;
    *out = nullptr;
    if (k < 1 || k > 32 || L < 1 || L > 10 || nnodes < 2 || !parent || !desc || !weight || !is_leaf)
        return plf_fail(ctx, PLF_ERR_INVALID, "plf_vocab_create: bad arguments (branching factor must be 1..32)");
    // children in insertion (= node id) order and word ids in leaf order, like loadFromTextFile (:1377-1418)
    std::vector<int> cnt(nnodes + 1, 0), word(nnodes, -1);
    for (int i = 1; i < nnodes; i++) {
        if (parent[i] < 0 || parent[i] >= i) return plf_fail(ctx, PLF_ERR_INVALID, "node %d: parent %d must precede it", i, parent[i]);
        cnt[parent[i] + 1]++;
    }
    for (int i = 0; i < nnodes; i++) {
        if (cnt[i + 1] > 32) return plf_fail(ctx, PLF_ERR_INVALID, "node %d has more than 32 children", i);
        cnt[i + 1] += cnt[i];
    }
    std::vector<int> ids(nnodes > 1 ? nnodes - 1 : 1), fill(cnt.begin(), cnt.end() - 1);
    for (int i = 1; i < nnodes; i++) ids[fill[parent[i]]++] = i;
    int nwords = 0;
    for (int i = 1; i < nnodes; i++)
        if (is_leaf[i]) {
            if (cnt[i + 1] != cnt[i]) return plf_fail(ctx, PLF_ERR_INVALID, "node %d is marked as a word but has children", i);
            word[i] = nwords++;
        }
    plf_vocab* v = new plf_vocab();
    v->ctx = ctx; v->k = k; v->L = L; v->nnodes = nnodes; v->nwords = nwords; v->scoring = scoring; v->weighting = weighting;
    v->parent.assign(parent, parent + nnodes);
    v->d_child_off = nullptr; v->d_child_ids = nullptr; v->d_desc = nullptr; v->d_word = nullptr; v->d_weight = nullptr;
    cudaSetDevice(ctx->device);
    cudaError_t e = cudaMalloc((void**)&v->d_child_off, (size_t)(nnodes + 1) * 4);
    if (e == cudaSuccess) e = cudaMalloc((void**)&v->d_child_ids, (size_t)nnodes * 4);
    if (e == cudaSuccess) e = cudaMalloc((void**)&v->d_desc, (size_t)nnodes * 32);
    if (e == cudaSuccess) e = cudaMalloc((void**)&v->d_word, (size_t)nnodes * 4);
    if (e == cudaSuccess) e = cudaMalloc((void**)&v->d_weight, (size_t)nnodes * 8);
    if (e == cudaSuccess) e = cudaMemcpy(v->d_child_off, cnt.data(), (size_t)(nnodes + 1) * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(v->d_child_ids, ids.data(), (size_t)(nnodes - 1) * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(v->d_desc, desc, (size_t)nnodes * 32, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(v->d_word, word.data(), (size_t)nnodes * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(v->d_weight, weight, (size_t)nnodes * 8, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        plf_status st = plf_fail(ctx, PLF_ERR_CUDA, "vocabulary upload failed: %s", cudaGetErrorString(e));
        plf_vocab_destroy(v);
        return st;
    }
    *out = v;
    return PLF_OK;
}

// ORBvoc.txt format of TemplatedVocabulary::loadFromTextFile (:1338-1424): "k L scoring weighting", then one line per
// node: parent isLeaf d0 .. d31 weight
extern "C" plf_status plf_vocab_load_text(plf_ctx* ctx, const char* path, plf_vocab** out)
{
    if (!ctx || !path || !out) return PLF_ERR_INVALID;
    std::ifstream f(path);
    if (!f.good()) return plf_fail(ctx, PLF_ERR_INVALID, "cannot open vocabulary file %s", path);
    std::string s;
    std::getline(f, s);
    std::stringstream ss(s);
    int k = -1, L = -1, n1 = -1, n2 = -1;
    ss >> k >> L >> n1 >> n2;
    if (k < 0 || k > 20 || L < 1 || L > 10 || n1 < 0 || n1 > 5 || n2 < 0 || n2 > 3)
        return plf_fail(ctx, PLF_ERR_INVALID, "Vocabulary loading failure: This is not a correct text file!");
    std::vector<int> parent(1, 0);
    std::vector<uint8_t> desc(32, 0), leaf(1, 0);
    std::vector<double> weight(1, 0.0);
    while (std::getline(f, s)) {
        if (s.empty()) continue;
        std::stringstream sn(s);
        int pid = 0, isleaf = 0;
        sn >> pid >> isleaf;
        if (sn.fail()) continue;
        parent.push_back(pid);
        leaf.push_back(isleaf > 0);
        for (int i = 0; i < 32; i++) { int v = 0; sn >> v; desc.push_back(sn.fail() ? 0 : (uint8_t)v); }
        double w = 0;
        sn >> w;
        weight.push_back(w);
    }
    return plf_vocab_create(ctx, k, L, n1, n2, (int)parent.size(), parent.data(), desc.data(), weight.data(), leaf.data(), out);
}

extern "C" void plf_vocab_destroy(plf_vocab* v)
{
    if (!v) return;
    cudaSetDevice(v->ctx->device);
    if (v->d_child_off) cudaFree(v->d_child_off);
    if (v->d_child_ids) cudaFree(v->d_child_ids);
    if (v->d_desc) cudaFree(v->d_desc);
    if (v->d_word) cudaFree(v->d_word);
    if (v->d_weight) cudaFree(v->d_weight);
    delete v;
}

extern "C" plf_status plf_vocab_info(const plf_vocab* v, int* k, int* L, int* nnodes, int* nwords, int* scoring, int* weighting)
{
    if (!v) return PLF_ERR_INVALID;
    if (k) *k = v->k;
    if (L) *L = v->L;
    if (nnodes) *nnodes = v->nnodes;
    if (nwords) *nwords = v->nwords;
    if (scoring) *scoring = v->scoring;
    if (weighting) *weighting = v->weighting;
    return PLF_OK;
}

extern "C" plf_status plf_bow_transform_device(plf_vocab* v, const uint8_t* dev_desc, int n, int levelsup, int32_t* dev_word,
                                               double* dev_weight, int32_t* dev_node)
{
    if (!v) return PLF_ERR_INVALID;
    plf_ctx* ctx = v->ctx;
    if (n < 0 || (n > 0 && (!dev_desc || !dev_word || !dev_weight || !dev_node))) return plf_fail(ctx, PLF_ERR_INVALID, "plf_bow_transform_device: bad arguments");
    if (((uintptr_t)dev_desc) & 15) return plf_fail(ctx, PLF_ERR_INVALID, "descriptor pointer must be 16-byte aligned");
    if (n == 0) return PLF_OK;
    PLF_CUDA(ctx, cudaSetDevice(ctx->device));
    PLF_LAUNCH(k_bow_transform, dim3(plf_div_up(n, 8)), dim3(256), 0, ctx->stream, (const uint4*)dev_desc, n, (const int*)v->d_child_off,
               (const int*)v->d_child_ids, (const uint4*)v->d_desc, (const int*)v->d_word, (const double*)v->d_weight, v->L - levelsup, dev_word,
               dev_weight, dev_node);
    PLF_CHECK_LAUNCH(ctx);
    return PLF_OK;
}

extern "C" plf_status plf_bow_transform(plf_vocab* v, const uint8_t* host_desc, int n, int levelsup, int32_t* host_word, double* host_weight,
                                        int32_t* host_node)
{
    if (!v) return PLF_ERR_INVALID;
    plf_ctx* ctx = v->ctx;
    if (n < 0 || (n > 0 && (!host_desc || !host_word || !host_weight || !host_node))) return plf_fail(ctx, PLF_ERR_INVALID, "plf_bow_transform: bad arguments");
    if (n == 0) return PLF_OK;
    PLF_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t db = plf_align_up((size_t)n * 32, 256), ib = plf_align_up((size_t)n * 4, 256), wb = plf_align_up((size_t)n * 8, 256);
    void* s;
    plf_status st = plf_ctx_scratch(ctx, db + 2 * ib + wb, &s);
    if (st) return st;
    uint8_t* p = (uint8_t*)s;
    uint8_t* dd = p; p += db;
    double* dw = (double*)p; p += wb;
    int* dword = (int*)p; p += ib;
    int* dnode = (int*)p;
    cudaStream_t sq = ctx->stream;
    PLF_CUDA(ctx, cudaMemcpyAsync(dd, host_desc, (size_t)n * 32, cudaMemcpyHostToDevice, sq));
    st = plf_bow_transform_device(v, dd, n, levelsup, dword, dw, dnode);
    if (st) return st;
    PLF_CUDA(ctx, cudaMemcpyAsync(host_word, dword, (size_t)n * 4, cudaMemcpyDeviceToHost, sq));
    PLF_CUDA(ctx, cudaMemcpyAsync(host_weight, dw, (size_t)n * 8, cudaMemcpyDeviceToHost, sq));
    PLF_CUDA(ctx, cudaMemcpyAsync(host_node, dnode, (size_t)n * 4, cudaMemcpyDeviceToHost, sq));
    PLF_CUDA(ctx, cudaStreamSynchronize(sq));
    return PLF_OK;
}
