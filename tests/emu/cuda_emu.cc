// cuda_emu.cc -- runtime for the TEST-ONLY CUDA emulation (see cuda_emu.h).
#include "cuda_emu.h"
#include <mutex>
#include <condition_variable>

thread_local uint3_ threadIdx, blockIdx;
thread_local dim3 blockDim, gridDim;
thread_local int emu_lane_linear;
thread_local unsigned char* emu_dyn_smem_ptr;

namespace {

struct Barrier {
    std::mutex m;
    std::condition_variable cv;
    int count = 0, waiting = 0;
    unsigned gen = 0;
    void reset(int n) { count = n; waiting = 0; }
    void wait()
    {
        std::unique_lock<std::mutex> lk(m);
        if (++waiting >= count) { waiting = 0; gen++; cv.notify_all(); return; }
        unsigned g = gen;
        cv.wait(lk, [&] { return g != gen; });
    }
    void leave()
    {
        std::unique_lock<std::mutex> lk(m);
        count--;
        if (count > 0 && waiting >= count) { waiting = 0; gen++; cv.notify_all(); }
    }
};

struct WarpState {
    Barrier bar;
    unsigned long long slot[32];
    unsigned alive = 0;
    std::mutex m;
};

const int MAXT = 1024;
// heap objects that are never destroyed: worker threads still wait on them at process exit
Barrier& g_block_bar = *new Barrier;
WarpState* g_warps = new WarpState[MAXT / 32];

struct Pool {
    std::mutex m;
    std::condition_variable cv_warp[1024 / 32], cv_done;      // launches wake only the warps they use
    std::vector<pthread_t> threads;
    unsigned long long epoch = 0;
    int nthreads_active = 0, ndone = 0;
    const std::function<void()>* body = nullptr;
    dim3 grid, block;
    uint3_ bidx;
    unsigned char* dyn = nullptr;
};
Pool& g_pool = *new Pool;

thread_local int t_warp;

void* worker(void* arg)
{
    int id = (int)(intptr_t)arg;
    unsigned long long seen = 0;
    for (;;) {
        {
            std::unique_lock<std::mutex> lk(g_pool.m);
            g_pool.cv_warp[id / 32].wait(lk, [&] { return g_pool.epoch != seen; });
            seen = g_pool.epoch;
            if (id >= g_pool.nthreads_active) continue;
        }
        dim3 b = g_pool.block;
        blockDim = b; gridDim = g_pool.grid; blockIdx = g_pool.bidx;
        threadIdx.x = id % b.x; threadIdx.y = (id / b.x) % b.y; threadIdx.z = id / (b.x * b.y);
        emu_lane_linear = id;
        emu_dyn_smem_ptr = g_pool.dyn;
        t_warp = id / 32;
        (*g_pool.body)();
        // thread exit: leave block barrier and warp
        {
            WarpState& w = g_warps[t_warp];
            std::unique_lock<std::mutex> lk(w.m);
            w.alive &= ~(1u << (id & 31));
        }
        g_block_bar.leave();
        {
            std::unique_lock<std::mutex> lk(g_pool.m);
            if (++g_pool.ndone == g_pool.nthreads_active) g_pool.cv_done.notify_all();
        }
    }
    return nullptr;
}

void ensure_threads(int n)
{
    while ((int)g_pool.threads.size() < n) {
        pthread_t t;
        pthread_attr_t a;
        pthread_attr_init(&a);
        pthread_attr_setstacksize(&a, 1 << 20);
        pthread_create(&t, &a, worker, (void*)(intptr_t)g_pool.threads.size());
        pthread_attr_destroy(&a);
        g_pool.threads.push_back(t);
    }
}

std::mutex& g_launch_mutex = *new std::mutex;

}  // namespace

void emu_launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body)
{
    std::lock_guard<std::mutex> launch_lock(g_launch_mutex);
    int nt = (int)(block.x * block.y * block.z);
    if (nt <= 0 || nt > MAXT) { fprintf(stderr, "emu: bad block size %d\n", nt); abort(); }
    ensure_threads(nt);
    std::vector<unsigned char> dyn(smem + 16);
    for (unsigned bz = 0; bz < grid.z; bz++)
        for (unsigned by = 0; by < grid.y; by++)
            for (unsigned bx = 0; bx < grid.x; bx++) {
                g_block_bar.reset(nt);
                for (int w = 0; w * 32 < nt; w++) {
                    int n = std::min(32, nt - w * 32);
                    g_warps[w].alive = n == 32 ? 0xffffffffu : ((1u << n) - 1);
                }
                {
                    std::unique_lock<std::mutex> lk(g_pool.m);
                    g_pool.body = &body; g_pool.grid = grid; g_pool.block = block;
                    g_pool.bidx.x = bx; g_pool.bidx.y = by; g_pool.bidx.z = bz;
                    g_pool.dyn = dyn.data();
                    g_pool.nthreads_active = nt; g_pool.ndone = 0;
                    g_pool.epoch++;
                    for (int w = 0; w * 32 < nt; w++) g_pool.cv_warp[w].notify_all();
                    g_pool.cv_done.wait(lk, [&] { return g_pool.ndone == nt; });
                }
            }
}

void emu_syncthreads() { g_block_bar.wait(); }

static void warp_barrier(WarpState& w, unsigned mask)
{
    unsigned m;
    {
        std::unique_lock<std::mutex> lk(w.m);
        m = mask & w.alive;
    }
    int n = __builtin_popcount(m);
    if (n <= 1) return;
    {
        std::unique_lock<std::mutex> lk(w.bar.m);
        w.bar.count = n;
    }
    w.bar.wait();
}

void emu_syncwarp(unsigned mask) { warp_barrier(g_warps[t_warp], mask); }

unsigned emu_ballot(unsigned mask, int pred)
{
    WarpState& w = g_warps[t_warp];
    int lane = emu_lane_linear & 31;
    w.slot[lane] = pred ? 1 : 0;
    warp_barrier(w, mask);
    unsigned alive;
    {
        std::unique_lock<std::mutex> lk(w.m);
        alive = w.alive;
    }
    unsigned r = 0;
    for (int i = 0; i < 32; i++)
        if (((mask & alive) >> i) & 1u) r |= (unsigned)(w.slot[i] & 1) << i;
    warp_barrier(w, mask);
    return r;
}

unsigned long long emu_shfl64(unsigned mask, unsigned long long v, int srcLane)
{
    WarpState& w = g_warps[t_warp];
    int lane = emu_lane_linear & 31;
    w.slot[lane] = v;
    warp_barrier(w, mask);
    unsigned long long r = w.slot[srcLane & 31];
    warp_barrier(w, mask);
    return r;
}
