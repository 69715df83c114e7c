"""CPU, world_size 2 over gloo: the N > 1 host logic (frame sharding, train sharding + top-2 merge).
The merge runs through the emulated kernels (tests/emu) because there is no GPU here; on the GPU box the
same code path runs over NCCL (tests/test_gpu_parity.py covers the merge kernel itself)."""
import os
import sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, emu_lib, q, t, out):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import spl_slam_b200 as S
    from spl_slam_b200 import sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ctx = S.Context(0, emu_lib)
    b, e = sharded.train_shard(len(t), rank, world)
    idx, dst = sharded.knn2_sharded(ctx, torch.from_numpy(q), torch.from_numpy(t[b:e].copy()), b)
    m12, nm = sharded.nnr_from_knn2(ctx, idx, dst, 0.75)
    np.savez(out % rank, idx=idx.numpy(), dist=dst.numpy(), m12=m12.numpy(), nm=nm, frames=np.array(sharded.frame_shard(11, rank, world)))
    dist.destroy_process_group()


def test_train_and_frame_shards_cover_everything():
    from spl_slam_b200 import sharded
    for nt in (0, 1, 7, 1000, 10007):
        for world in (1, 2, 3, 8):
            rows = []
            for r in range(world):
                b, e = sharded.train_shard(nt, r, world)
                assert all(j * world // nt == r for j in range(b, e)) if nt else b == e == 0
                rows += list(range(b, e))
            assert rows == list(range(nt))
    for world in (1, 2, 4, 8):
        got = sorted(sum((sharded.frame_shard(37, r, world) for r in range(world)), []))
        assert got == list(range(37))
    assert sharded.frame_shard(8, 0, 2) == [0, 2, 4, 6] and sharded.frame_shard(8, 1, 2) == [1, 3, 5, 7]   # left / right


def test_sharded_knn2_world2_gloo(oracle, emu_lib, tmp_path):
    import torch.multiprocessing as mp
    rng = np.random.default_rng(7)
    q = rng.integers(0, 4, (150, 32), dtype=np.uint8)     # tie-heavy: the merge must keep the lowest index
    t = rng.integers(0, 4, (701, 32), dtype=np.uint8)
    out = str(tmp_path / "r%d.npz")
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, emu_lib, q, t, out), nprocs=2, join=True)
    oi, od = oracle.knn2(q, t)
    om, on = oracle.match_nnr(q, t, 0.75)
    for r in range(2):
        z = np.load(out % r)
        assert np.array_equal(z["idx"], oi) and np.array_equal(z["dist"], od)
        assert np.array_equal(z["m12"], om) and int(z["nm"]) == on
    assert list(np.load(out % 0)["frames"]) == [0, 2, 4, 6, 8, 10]
