"""The row-tiled path behind the C ABI (fd_tiled_*, csrc/fd_tiled.cu): one process, one context per tile, halo rows by device-to-device
copies, candidate keys packed on the first tile's device by a kernel reading the other tiles' slots, selection there.  With one GPU
all tiles live on it (ordinals repeat) -- the seam, halo, gather and selection code is the same; with several GPUs the tiles are
spread over them (gpurun --gpus N, and bench.py's parity_checks at world > 1).

Reference: the dense stages of feature_point_harris_detector.cpp:17-137 / feature_point_fast_detector.cpp:83-98 over the WHOLE frame and
feature_point_detector.cpp:54-74 -- a tiled run has to return exactly what the untiled run returns."""
import numpy as np
import pytest

import feature_detector_b200 as fd
from feature_detector_b200.synth import synth

pytestmark = pytest.mark.gpu

CASES = [(fd.HARRIS, 30.0, 20, 200, 12), (fd.SHI_TOMAS, 40.0, 20, 1000, 12), (fd.FAST, 10.0, 20, 200, 9), (fd.FAST, 10.0, 20, 200, 12),
         (fd.FAST, 0.1, 15, 300, 12)]


def _devices(n_tiles):
    import torch
    g = max(torch.cuda.device_count(), 1)
    return [k % g for k in range(n_tiles)]


def _untiled(frames, prm, n):
    with fd.Context(0) as ctx:
        ctx.upload(frames)
        ctx.detect(prm)
        kp, cnt = ctx.keypoints(max(n, 1))
        cands = [ctx.candidates(f) for f in range(len(frames))]
    return kp, cnt, cands


@pytest.mark.parametrize("shape", [(333, 217, 3), (752, 480, 2), (160, 40, 1), (1001, 37, 2)])
@pytest.mark.parametrize("n_tiles", [1, 2, 3, 8])
def test_tiled_equals_untiled(shape, n_tiles):
    w, h, nf = shape
    frames = np.stack([synth(w, h, 40 + i) for i in range(nf)])
    with fd.TiledDetector(_devices(n_tiles)) as td:
        td.upload(frames)
        for kind, thr, d, n, fast_n in CASES:
            prm = fd.DetectParams(kind, thr, d, n, fast_n=fast_n)
            kp_ref, cnt_ref, cand_ref = _untiled(frames, prm, n)
            td.exchange_halos()          # a second exchange must leave the halos as they are
            td.detect(prm)
            kp, cnt = td.keypoints(max(n, 1))
            assert np.array_equal(cnt, cnt_ref), (kind, thr, n_tiles)
            assert np.array_equal(td.candidate_counts(), [len(c) for c in cand_ref])
            for f in range(nf):
                c = td.candidates(f)
                assert np.array_equal(c, cand_ref[f]), (kind, thr, n_tiles, f)            # seam-free candidates, FAST offsets included
                assert np.array_equal(kp[f, :cnt[f]], kp_ref[f, :cnt[f]]), (kind, thr, n_tiles, f)


def test_more_tiles_than_rows_and_tiny_frames():
    rng = np.random.default_rng(5)
    for rows, cols in ((5, 40), (9, 131), (2, 2), (16, 16)):
        frames = rng.integers(0, 256, (2, rows, cols), dtype=np.uint8)
        prm = fd.DetectParams(fd.HARRIS, 0.1, 2, 50)
        kp_ref, cnt_ref, cand_ref = _untiled(frames, prm, 50)
        with fd.TiledDetector(_devices(7)) as td:
            td.upload(frames)
            td.detect(prm)
            kp, cnt = td.keypoints(50)
            assert np.array_equal(cnt, cnt_ref), (rows, cols)
            for f in range(2):
                assert np.array_equal(td.candidates(f), cand_ref[f]) and np.array_equal(kp[f, :cnt[f]], kp_ref[f, :cnt[f]])


def test_scatter_from_device_frames_and_in_place_producer():
    """Frames resident on one device are split by peer copies; afterwards a producer rewrites the tiles' own rows in place
    (fd_tiled_tile_info) and only the halos travel again."""
    import torch
    frames = np.stack([synth(640, 360, 7 + i) for i in range(3)])
    newer = np.stack([synth(640, 360, 90 + i) for i in range(3)])
    prm = fd.DetectParams(fd.HARRIS, 30.0, 20, 200)
    d_frames = torch.from_numpy(frames).cuda()
    with fd.TiledDetector(_devices(4)) as td:
        td.scatter_device(d_frames.data_ptr(), 360, 640, 3)
        td.detect(prm)
        kp, cnt = td.keypoints(200)
        kp_ref, cnt_ref, _ = _untiled(frames, prm, 200)
        assert np.array_equal(cnt, cnt_ref) and all(np.array_equal(kp[f, :cnt[f]], kp_ref[f, :cnt[f]]) for f in range(3))
        assert td.halo_bytes == 3 * 2 * 3 * 3 * 640      # three interior seams, two directions, three rows, three frames
        td.sync()
        for k in range(4):                                # the producer: new own rows, written on the tile's device
            info = td.tile_info(k)
            lo, cnt_rows = info["own_first_row"], info["own_row_count"]
            with torch.cuda.device(info["device"]):
                src = torch.from_numpy(newer[:, lo:lo + cnt_rows].copy()).cuda()
                from cuda.bindings import runtime as cudart
                for f in range(3):
                    err, = cudart.cudaMemcpy2D(info["ptr"] + f * info["frame_stride"], info["pitch"], src[f].data_ptr(), 640, 640, cnt_rows,
                                               cudart.cudaMemcpyKind.cudaMemcpyDeviceToDevice)
                    assert int(err) == 0
        td.exchange_halos()
        td.detect(prm)
        kp, cnt = td.keypoints(200)
        kp_ref, cnt_ref, _ = _untiled(newer, prm, 200)
        assert np.array_equal(cnt, cnt_ref) and all(np.array_equal(kp[f, :cnt[f]], kp_ref[f, :cnt[f]]) for f in range(3))


def test_tile_capacity_overflow_is_reported():
    frames = synth(752, 480, 1)[None]
    with fd.TiledDetector(_devices(2)) as td:
        td.upload(frames)
        td.detect(fd.DetectParams(fd.FAST, 0.1, 15, 200), 1000)
        with pytest.raises(fd.FdError):
            td.keypoints(200)
        td.detect(fd.DetectParams(fd.FAST, 0.1, 15, 200), 0)    # and the detector is usable afterwards
        assert td.keypoints(200)[1][0] == 200


def test_config3_frame_tiled_over_all_gpus():
    """BASELINE.json configs[3]: one 3840x2160 Harris frame over every GPU of the box (one GPU: 4 tiles on it)."""
    import torch
    g = torch.cuda.device_count()
    frame = synth(3840, 2160, 0)[None]
    prm = fd.DetectParams(fd.HARRIS, 30.0, 20, 200)
    kp_ref, cnt_ref, cand_ref = _untiled(frame, prm, 200)
    with fd.TiledDetector(list(range(g)) if g > 1 else [0, 0, 0, 0]) as td:
        td.upload(frame)
        td.detect(prm, 1 << 20)
        kp, cnt = td.keypoints(200)
        assert cnt[0] == cnt_ref[0] == 200 and np.array_equal(kp[0, :200], kp_ref[0, :200])
        assert np.array_equal(td.candidates(0), cand_ref[0])


def test_first_range_prefilter_equals_full_gather(monkeypatch):
    """fd_tiled_detect ships only the keys below each frame's first rank limit to the root (tile histograms summed there, limits handed
    back, frames that need more flagged on the device and finished from a full gather); FD_B200_TILED_PREFILTER=0 gathers every key as
    before.  Same keypoints and candidates either way: frames with many candidates (the prefiltered way), with few (flagged), FAST at the
    reference's default threshold (a first range larger than the gathered slots), needed counts that one range cannot satisfy."""
    frames = np.stack([synth(752, 480, 60 + i) for i in range(3)])
    cases = [(fd.HARRIS, 30.0, 20, 200, 12), (fd.HARRIS, 0.1, 3, 4000, 12), (fd.FAST, 0.1, 15, 200, 12), (fd.FAST, 10.0, 20, 200, 9),
             (fd.SHI_TOMAS, 40.0, 20, 1000, 12), (fd.HARRIS, 30.0, 20, 0, 12), (fd.HARRIS, 1e9, 20, 50, 12)]
    outs = []
    for flag in ("0", "1"):
        monkeypatch.setenv("FD_B200_TILED_PREFILTER", flag)
        with fd.TiledDetector(_devices(3)) as td:
            td.upload(frames)
            res = []
            for kind, thr, d, n, fast_n in cases:
                prm = fd.DetectParams(kind, thr, d, n, fast_n=fast_n)
                td.detect(prm)
                kp, cnt = td.keypoints(max(n, 1))
                res.append((kp.copy(), cnt.copy(), td.candidate_counts(), [td.candidates(f) for f in range(len(frames))]))
                td.detect(prm)                                     # and again on warm buffers
                kp2, cnt2 = td.keypoints(max(n, 1))
                assert np.array_equal(cnt, cnt2) and np.array_equal(kp, kp2)
            outs.append(res)
    for (kp0, cnt0, cc0, cand0), (kp1, cnt1, cc1, cand1), case in zip(outs[0], outs[1], cases):
        assert np.array_equal(cnt0, cnt1) and np.array_equal(cc0, cc1), case
        for f in range(len(frames)):
            assert np.array_equal(kp0[f, :cnt0[f]], kp1[f, :cnt1[f]]) and np.array_equal(cand0[f], cand1[f]), (case, f)
    kp_ref, cnt_ref, _ = _untiled(frames, fd.DetectParams(fd.HARRIS, 30.0, 20, 200), 200)
    assert np.array_equal(outs[1][0][1], cnt_ref) and all(np.array_equal(outs[1][0][0][f, :cnt_ref[f]], kp_ref[f, :cnt_ref[f]]) for f in range(3))
