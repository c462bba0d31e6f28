"""A batch of frames sharded across ranks (first row): block partition, the optional gather of the fixed-capacity
result slots -- over gloo on CPU with the oracle port as the per-frame kernel, over NCCL on however many GPUs the box has."""
import os
import socket

import numpy as np
import pytest

from feature_detector_b200 import sharding
from feature_detector_b200.synth import synth

W, H, N_FRAMES, NEEDED = 160, 120, 7, 40


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_frame_range_partitions_the_batch():
    for n in (0, 1, 7, 8, 1024, 4096):
        for world in (1, 2, 3, 8):
            blocks = [sharding.frame_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[r][1] == blocks[r + 1][0] for r in range(world - 1))          # contiguous, in rank order
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)    # longer blocks first
    assert sharding.frame_range(4096, 3, 8) == (1536, 2048)                                  # configs[2]: 512 frames per GPU


def _port_slots(frames):
    """Per-frame result slots as the GPU path lays them out, computed by the oracle port (FAST kN 9 thr 10 d 20 + BRIEF-256)."""
    from oracle.bindings import FAST, Port
    port = Port()
    kp = np.zeros((len(frames), NEEDED, 4), np.float32)
    cnt = np.zeros(len(frames), np.int32)
    desc = np.zeros((len(frames), NEEDED, 32), np.uint8)
    for f, im in enumerate(frames):
        o = port.detect(FAST, im, 10.0, 20, NEEDED, fast_n=9)
        n = len(o["features"])
        kp[f, :n, :2] = o["features"]
        cnt[f] = n
        _, bits = port.brief(im, o["features"], 256, 8)
        desc[f, :n] = np.packbits(bits.astype(np.uint8), axis=-1, bitorder="little")
    return kp, cnt, desc


def _gloo_worker(rank, world, port, ret):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = sharding.frame_range(N_FRAMES, rank, world)
        mine = [synth(W, H, i) for i in range(lo, hi)]
        kp, cnt, desc = _port_slots(mine)
        got = [sharding.gather_frame_slots(torch.from_numpy(a), N_FRAMES, rank, world, dst=0) for a in (kp, cnt, desc)]
        everywhere = sharding.gather_frame_slots(torch.from_numpy(cnt), N_FRAMES, rank, world)   # no dst: every rank gets it
        ret[f"all{rank}"] = everywhere.numpy().tolist()
        # the packed form: all three arrays in one all_gather
        packed = sharding.gather_frame_records([torch.from_numpy(a) for a in (kp, cnt, desc)], N_FRAMES, rank, world, dst=0)
        packed_all = sharding.gather_frame_records([torch.from_numpy(cnt)], N_FRAMES, rank, world)
        ret[f"packed_all{rank}"] = packed_all[0].numpy().tolist()
        if rank == 0:
            want = _port_slots([synth(W, H, i) for i in range(N_FRAMES)])
            ret["same"] = all(np.array_equal(g.numpy(), w) for g, w in zip(got, want))
            ret["same_packed"] = all(g.dtype == torch.from_numpy(w).dtype and np.array_equal(g.numpy(), w) for g, w in zip(packed, want))
            ret["counts"] = want[1].tolist()
        else:
            ret[f"none{rank}"] = all(g is None for g in got) and packed is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3, 8])      # 8 ranks for 7 frames: the last block is empty
def test_gathered_slots_equal_the_unsharded_batch_over_gloo(world, built):
    import torch.multiprocessing as mp
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_gloo_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
        assert ret["same"] and sum(ret["counts"]) > 0, dict(ret)
        assert all(ret[f"none{r}"] for r in range(1, world))
        assert all(ret[f"all{r}"] == ret["counts"] for r in range(world))
        assert ret["same_packed"] and all(ret[f"packed_all{r}"] == ret["counts"] for r in range(world))


def _nccl_worker(rank, world, port, ret):
    import torch
    import torch.distributed as dist
    import feature_detector_b200 as fd
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        n = 5 * world + 3
        lo, hi = sharding.frame_range(n, rank, world)
        det, brief = fd.DetectParams(fd.FAST, 10.0, 20, NEEDED, fast_n=9), fd.BriefParams(256, 8)
        block = np.stack([synth(W, H, i) for i in range(lo, hi)])
        with fd.Context(rank) as ctx:
            src = torch.from_numpy(block).cuda() if rank % 2 == 0 else block        # bound device frames and uploaded host frames
            got = sharding.detect_sharded(ctx, src, n, rank, world, det, brief, gather=True, dst=0)
            local = sharding.detect_sharded(ctx, src, n, rank, world, det, brief, gather=False)
            if rank == 0:
                ctx.upload(np.stack([synth(W, H, i) for i in range(n)]))
                ctx.detect(det)
                ctx.describe_selected(brief)
                kp, cnt = ctx.keypoints(NEEDED)
                desc = ctx.descriptors(NEEDED)
                same = np.array_equal(got["counts"], cnt)
                for f in range(n):
                    same &= np.array_equal(got["keypoints"][f, :cnt[f]], kp[f, :cnt[f]])
                    same &= np.array_equal(got["descriptors"][f, :cnt[f]], desc[f, :cnt[f]])
                same &= np.array_equal(local["counts"], cnt[lo:hi]) and np.array_equal(local["keypoints"][0, :cnt[lo]], kp[lo, :cnt[lo]])
                ret["same"] = bool(same)
                ret["n_kp"] = int(cnt.sum())
            else:
                ret[f"none{rank}"] = got is None
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_sharded_batch_with_gather_over_nccl():
    import torch
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 8)       # one GPU: the gather degenerates, the device-pointer views are still exercised
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_nccl_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
        assert ret["same"] and ret["n_kp"] > 0, dict(ret)
        assert all(ret[f"none{r}"] for r in range(1, world))
