"""Worker of tests/test_sharding.py::test_gloo_world_size_2 (one process per rank, gloo backend, CPU)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lio_slam_b200 import sharding, synth  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402


def main():
    dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
    rank, ws = dist.get_rank(), dist.get_world_size()
    o = Oracle("port")
    world = synth.make_world(1234)
    clouds = []
    for i in range(4):
        p = synth.path_pose(-0.8 * i)
        clouds.append(synth.transform_packed(synth.to_packed(synth.make_scan(world, p, 16, seed=400 + i, cols=300)), p))
    cloud = np.concatenate(clouds)
    tile, _ = sharding.plan_voxel_tiles(cloud, 0.5, ws)          # every rank derives the same plan
    mine, _ = o.voxel_grid(sharding.shard_points(cloud, tile, rank), 0.5)
    # ordered gather: sizes first, then padded payloads (no collective is needed on a real data path —
    # each GPU writes its slice of the host buffer — this only exercises the N>1 plumbing)
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(ws)]
    dist.all_gather(sizes, torch.tensor([mine.shape[0]], dtype=torch.int64))
    cap = int(max(s.item() for s in sizes))
    pad = torch.zeros((cap, 4), dtype=torch.float32)
    pad[: mine.shape[0]] = torch.from_numpy(mine)
    bufs = [torch.zeros((cap, 4), dtype=torch.float32) for _ in range(ws)]
    dist.all_gather(bufs, pad)
    seqs = sharding.assign_sequences(8, ws, rank)
    cnt = torch.tensor([len(seqs)], dtype=torch.int64)
    dist.all_reduce(cnt)
    if rank == 0:
        got = np.concatenate([bufs[r][: int(sizes[r].item())].numpy() for r in range(ws)])
        want, _ = o.voxel_grid(cloud, 0.5)
        assert got.shape == want.shape and np.array_equal(got.view(np.uint32), want.view(np.uint32))
        assert int(cnt.item()) == 8
        print("GLOO_OK", got.shape[0])
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
