"""Multi-GPU sharding of the hot path (one process per GPU, torch.distributed for the plumbing).

Extraction shards by frame with no data-path collective (frame i -> rank i mod N; the left/right images of
a stereo pair land on neighbouring ranks).  Brute-force matching shards the TRAIN set by rows; every rank
computes its local top-2 with global indices, the per-shard tables are all-gathered (NCCL over NVLink on
GPUs, 16 B per query per rank) and merged by (distance, index), which reproduces the single-GPU
cv::BFMatcher ordering exactly, ties included.
"""
import numpy as np


def frame_shard(nframes, rank, world):
    """Global frame ids processed by `rank`."""
    return list(range(rank, nframes, world))


def train_shard(nt, rank, world):
    """Row range [begin, end) of the train set owned by `rank`: row j -> shard floor(j * world / nt)."""
    begin = -(-rank * nt // world)          # ceil(rank * nt / world)
    end = -(-(rank + 1) * nt // world)
    return begin, min(end, nt)


class Comm:
    """The C-ABI communicator (include/plf.h: plf_comm_*): one NCCL communicator owned by libplf.so.  Python only launches:
    rank 0's unique id travels through torch.distributed's broadcast, everything after that -- local top-2, ncclAllGather,
    merge, ratio test -- is queued by the library on the context stream."""

    def __init__(self, ctx, group=None):
        import torch
        import torch.distributed as dist
        import ctypes as C
        self.ctx = ctx
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = (C.c_uint8 * 128)()
            st = ctx.lib.plf_comm_unique_id(buf)
            if st:
                raise RuntimeError("plf_comm_unique_id failed (NCCL not loadable)")
            uid = torch.tensor(list(buf), dtype=torch.uint8)
        if world > 1:
            dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
            t = uid.to(dev)
            dist.broadcast(t, src=0, group=group)
            uid = t.cpu()
        arr = (C.c_uint8 * 128)(*uid.tolist())
        h = C.c_void_p()
        ctx.check(ctx.lib.plf_comm_create(ctx.h, arr, rank, world, C.byref(h)))
        self.h, self.rank, self.world = h, rank, world

    def close(self):
        if getattr(self, "h", None):
            self.ctx.lib.plf_comm_destroy(self.h)
            self.h = None

    def knn2(self, q, t_local, index_base):
        """Top-2 over the union of all ranks' shards (device tensors) -> (idx, dist) int32 (nq, 2), identical on every rank."""
        import torch
        nq = q.shape[0]
        idx = torch.empty((nq, 2), dtype=torch.int32, device=q.device)
        dst = torch.empty((nq, 2), dtype=torch.int32, device=q.device)
        self.ctx.check(self.ctx.lib.plf_hamming_knn2_sharded_device(self.ctx.h, self.h, q.data_ptr(), nq, t_local.data_ptr() if t_local.numel() else None,
                                                                    int(t_local.shape[0]), int(index_base), idx.data_ptr(), dst.data_ptr()))
        return idx, dst

    def match_nnr(self, q, t_local, index_base, nnr):
        """Linematcher::matchNNR over the sharded train set -> (idx, dist, matches12, nmatches tensor); asynchronous."""
        import torch
        nq = q.shape[0]
        idx = torch.empty((nq, 2), dtype=torch.int32, device=q.device)
        dst = torch.empty((nq, 2), dtype=torch.int32, device=q.device)
        m12 = torch.empty(nq, dtype=torch.int32, device=q.device)
        nm = torch.zeros(1, dtype=torch.int32, device=q.device)
        self.ctx.check(self.ctx.lib.plf_match_nnr_sharded_device(self.ctx.h, self.h, q.data_ptr(), nq, t_local.data_ptr() if t_local.numel() else None,
                                                                 int(t_local.shape[0]), int(index_base), float(nnr), idx.data_ptr(), dst.data_ptr(),
                                                                 m12.data_ptr(), nm.data_ptr()))
        return idx, dst, m12, nm


def knn2_sharded(ctx, q, t_local, index_base, group=None):
    """Top-2 of every query over the union of all ranks' train shards -- the torch.distributed formulation, kept for the
    CPU (gloo + emulated kernels) test of the N > 1 host logic; GPU runs use Comm (the C-ABI path) above.

    q: (nq, 32) uint8 torch tensor (replicated on every rank); t_local: this rank's (nt_local, 32) shard;
    index_base: global row index of t_local[0].  Tensors live where the context's library computes
    (CUDA for libplf.so).  Returns (idx, dist) int32 (nq, 2) tensors, identical on every rank.
    """
    import torch
    import torch.distributed as dist
    nq = q.shape[0]
    idx = torch.empty((nq, 2), dtype=torch.int32, device=q.device)
    dst = torch.empty((nq, 2), dtype=torch.int32, device=q.device)
    lib = ctx.lib
    ctx.check(lib.plf_hamming_knn2_device(ctx.h, q.data_ptr(), nq, t_local.data_ptr() if t_local.numel() else None,
                                          int(t_local.shape[0]), int(index_base), idx.data_ptr(), dst.data_ptr()))
    ctx.synchronize()
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return idx, dst
    # [shard][nq][2] contiguous, shaped as the dim-0 concatenation both gloo and NCCL accept
    pidx = torch.empty((world * nq, 2), dtype=torch.int32, device=q.device)
    pdst = torch.empty((world * nq, 2), dtype=torch.int32, device=q.device)
    dist.all_gather_into_tensor(pidx, idx, group=group)
    dist.all_gather_into_tensor(pdst, dst, group=group)
    if q.is_cuda:
        torch.cuda.current_stream().synchronize()
    ctx.check(lib.plf_knn2_merge_device(ctx.h, pidx.data_ptr(), pdst.data_ptr(), world, nq, idx.data_ptr(), dst.data_ptr()))
    ctx.synchronize()
    return idx, dst


def nnr_from_knn2(ctx, idx, dst, nnr):
    """Ratio test of Linematcher::matchNNR (src/Linematcher.cc:534-538) on a merged top-2 table."""
    import torch
    nq = idx.shape[0]
    m12 = torch.empty(nq, dtype=torch.int32, device=idx.device)
    nm = torch.zeros(1, dtype=torch.int32, device=idx.device)
    ctx.check(ctx.lib.plf_nnr_from_knn2_device(ctx.h, idx.data_ptr(), dst.data_ptr(), nq, float(nnr), m12.data_ptr(), nm.data_ptr()))
    ctx.synchronize()
    return m12, int(nm.item())
