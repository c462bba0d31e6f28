"""Row tiling of one large frame (BASELINE.json configs[3]): tiles + halo exchange + key gather + one selection.

CPU part (gloo, world_size 2 and 3): the distributed plumbing of feature_detector_b200/tiling.py, with the oracle standing
in for the per-tile kernel -- the tiled candidate set must equal the untiled one.
GPU part: the same tiles on ONE GPU through the C ABI (fd_set_tile / fd_export_candidates / fd_select_candidates), and,
where the box has at least two GPUs, the real multi-rank path over NCCL.
"""
import os
import socket

import numpy as np
import pytest

from feature_detector_b200 import tiling
from feature_detector_b200.synth import synth


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_plan_tiles_covers_the_frame():
    for rows, n in ((2160, 8), (480, 3), (10, 4), (5, 8), (1, 1)):
        tiles = tiling.plan_tiles(rows, n)
        assert len(tiles) == n and tiles[0].own_lo == 0 and tiles[-1].own_hi == rows
        for a, b in zip(tiles, tiles[1:]):
            assert a.own_hi == b.own_lo
        for t in tiles:
            if t.own_count:
                assert t.buf_lo == max(0, t.own_lo - 3) and t.buf_hi == min(rows, t.own_hi + 3)
                assert 0 <= t.own_first and t.own_first + t.own_count <= t.buf_rows


def test_key_codec_roundtrip_and_order():
    rng = np.random.default_rng(1)
    resp = np.concatenate([rng.normal(0, 1e4, 500).astype(np.float32), np.array([0.0, -0.0, 1e-30, 3.5, 3.5], np.float32)])
    rows = rng.integers(0, 65535, len(resp))
    cols = rng.integers(0, 65535, len(resp))
    keys = tiling.make_keys(resp, rows, cols)
    r2, y2, x2 = tiling.split_keys(keys)
    assert np.array_equal(r2.view(np.uint32), resp.view(np.uint32)) and np.array_equal(y2, rows) and np.array_equal(x2, cols)
    order = np.argsort(keys, kind="stable")
    assert np.all(np.diff(resp[order].astype(np.float64)) <= 0)          # ascending keys = descending response


def _gloo_worker(rank, world, port, rows, cols, idx, ret):
    import torch
    import torch.distributed as dist
    from oracle.bindings import HARRIS, Port
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        frame = synth(cols, rows, idx)
        tiles = tiling.plan_tiles(rows, world)
        me = tiles[rank]
        owned = torch.from_numpy(frame[me.own_lo:me.own_hi].copy())
        buf = tiling.exchange_halos(owned, rank, world, tiles)
        ok_halo = bool(np.array_equal(buf.numpy(), frame[me.buf_lo:me.buf_hi]))
        # the oracle plays the tile kernel: responses on the tile buffer, candidates of the own rows, absolute coordinates
        port_lib = Port()
        o = port_lib.detect(HARRIS, buf.numpy(), 30.0, 20, 1, want_candidates=True) if me.buf_rows >= 5 else {"cand_resp": np.zeros(0, np.float32), "cand_xy": np.zeros((0, 2), np.int32)}
        y = o["cand_xy"][:, 1] + me.buf_lo
        keep = (y >= me.own_lo) & (y < me.own_hi)
        keys = tiling.make_keys(o["cand_resp"][keep], y[keep], o["cand_xy"][keep, 0])
        allk = tiling.gather_keys(torch.from_numpy(keys.view(np.int64).copy()), rank, world)
        if rank == 0:
            full = port_lib.detect(HARRIS, frame, 30.0, 20, 1, want_candidates=True)
            want = np.sort(tiling.make_keys(full["cand_resp"], full["cand_xy"][:, 1], full["cand_xy"][:, 0]))
            got = np.sort(allk.numpy().view(np.uint64))
            ret["same"] = bool(np.array_equal(got, want))
            ret["n"] = int(len(want))
        ret[f"halo{rank}"] = ok_halo
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,rows,cols", [(2, 120, 160), (3, 97, 131), (2, 7, 64)])
def test_tiled_candidates_equal_untiled_over_gloo(world, rows, cols, built):
    import torch.multiprocessing as mp
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_gloo_worker, args=(world, _free_port(), rows, cols, 3, ret), nprocs=world, join=True)
        assert all(ret[f"halo{r}"] for r in range(world))
        assert ret["same"], dict(ret)


# ---- GPU ---------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("case", [("harris", 30.0, 20, 200, 12), ("shi", 40.0, 20, 300, 12), ("fast", 10.0, 20, 200, 9), ("fast", 10.0, 20, 200, 12),
                                  ("fast", 0.5, 8, 300, 9)])
@pytest.mark.parametrize("shape_tiles", [((752, 480, 2), 2), ((752, 480, 2), 8), ((333, 217, 5), 3), ((1280, 720, 1), 5), ((160, 40, 4), 7)])
def test_tiles_on_one_gpu_equal_the_untiled_run(case, shape_tiles):
    import torch
    import feature_detector_b200 as fd
    kinds = {"harris": fd.HARRIS, "shi": fd.SHI_TOMAS, "fast": fd.FAST}
    (w, h, idx), n_tiles = shape_tiles
    name, thr, d, n, fast_n = case
    im = synth(w, h, idx)
    prm = fd.DetectParams(kinds[name], thr, d, n, fast_n=fast_n)
    with fd.Context(0) as ctx:
        ctx.upload(im)
        ctx.detect(prm)
        kp_ref, cnt_ref = ctx.keypoints(max(n, 1))
        cand = ctx.candidates(0)
        want = np.sort(tiling.make_keys(cand["response"], cand["y"], cand["x"]))
        kp, keys = tiling.detect_tiled_local(ctx, torch.from_numpy(im).cuda(), n_tiles, prm)
        assert np.array_equal(np.sort(keys.cpu().numpy().view(np.uint64)), want)        # seam-free candidates (FAST offsets included)
        assert np.array_equal(kp, kp_ref[0, :cnt_ref[0]])


def _nccl_worker(rank, world, port, rows, cols, ret):
    import torch
    import torch.distributed as dist
    import feature_detector_b200 as fd
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        frame = synth(cols, rows, 0)
        me = tiling.plan_tiles(rows, world)[rank]
        owned = torch.from_numpy(frame[me.own_lo:me.own_hi].copy()).cuda()
        prm = fd.DetectParams(fd.HARRIS, 30.0, 20, 200)
        with fd.Context(rank) as ctx:
            kp = tiling.detect_tiled(ctx, owned, rows, prm, rank, world)
            if rank == 0:
                ctx.upload(frame)
                ctx.detect(prm)
                kp_ref, cnt = ctx.keypoints(200)
                ret["same"] = bool(np.array_equal(kp, kp_ref[0, :cnt[0]]))
                ret["n"] = int(cnt[0])
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_row_tiled_frame_over_nccl():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least two GPUs (run under gpurun --gpus 2)")
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 8)
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_nccl_worker, args=(world, _free_port(), 2160, 3840, ret), nprocs=world, join=True)
        assert ret["same"] and ret["n"] == 200
