"""BASELINE.json configs at their full frame shapes (GPU): sampled frames against the CPU checker plus size-independent
properties (a batch equals its frames processed one by one; duplicated frames give duplicated results; tiles equal the
untiled frame; sortedness of the LSD seed order)."""
import numpy as np
import pytest

import feature_detector_b200 as fd
from feature_detector_b200.synth import synth
from oracle.bindings import FAST, HARRIS, SHI_TOMAS

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = fd.Context(0)
    yield c
    c.close()


def _feats(kp, cnt, f):
    return np.stack([kp["x"][f, :cnt[f]], kp["y"][f, :cnt[f]]], 1).astype(np.float32)


def test_config1_fast_select_brief_752x480_batch(ctx, checker):
    """configs[1]: FAST (kN 9 and 12) + selection + BRIEF-256 on a 256-frame batch of 752x480 (64 distinct frames x 4)."""
    base = np.stack([synth(752, 480, 500 + i) for i in range(64)])
    frames = np.concatenate([base] * 4)
    ctx.upload(frames)
    for fast_n in (9, 12):
        ctx.detect(fd.DetectParams(fd.FAST, 10.0, 20, 200, fast_n=fast_n), 65536)
        ctx.describe_selected(fd.BriefParams(256, 8))
        kp, cnt = ctx.keypoints(200)
        desc = ctx.descriptors(200)
        cand = ctx.candidate_counts()
        # duplicated frames -> duplicated results (frames are independent)
        for rep in range(1, 4):
            assert np.array_equal(cnt[:64], cnt[64 * rep:64 * rep + 64]) and np.array_equal(cand[:64], cand[64 * rep:64 * rep + 64])
            assert np.array_equal(kp[:64], kp[64 * rep:64 * rep + 64]) and np.array_equal(desc[:64], desc[64 * rep:64 * rep + 64])
        for f in range(0, 64, 3 if fast_n == 9 else 7):     # 22 / 10 of the 64 distinct frames against the checker, frame by frame
            o = checker.detect(FAST, base[f], 10.0, 20, 200, fast_n=fast_n)
            assert cand[f] == o["n_cand"] and np.array_equal(_feats(kp, cnt, f), o["features"]), (fast_n, f)
            ok, bits = checker.brief(base[f], o["features"], 256, 8)
            assert np.array_equal(fd.unpack_bits(desc[f, :cnt[f]]), bits), (fast_n, f)


def test_config2_shi_tomasi_top1000_1280x720(ctx, checker):
    """configs[2]: Shi-Tomasi (the reference's larger-eigenvalue response), thr 40, d 20, N 1000 on 1280x720 frames."""
    from oracle.tiecheck import same_up_to_ties
    frames = np.stack([synth(1280, 720, 700 + i) for i in range(6)])
    ctx.upload(frames)
    ctx.detect(fd.DetectParams(fd.SHI_TOMAS, 40.0, 20, 1000), 0)
    kp, cnt = ctx.keypoints(1000)
    for f in range(6):
        o = checker.detect(SHI_TOMAS, frames[f], 40.0, 20, 1000)
        cand = ctx.candidates(f)
        assert len(cand) == o["n_cand"]
        g = np.lexsort((cand["x"], cand["y"]))
        c = np.lexsort((o["cand_xy"][:, 0], o["cand_xy"][:, 1]))
        assert np.array_equal(cand["response"][g].view(np.uint32), o["cand_resp"][c].view(np.uint32))   # responses bitwise
        if not np.array_equal(_feats(kp, cnt, f), o["features"]):
            assert same_up_to_ties(o["cand_resp"], o["cand_xy"], cand["response"], np.stack([cand["x"], cand["y"]], 1))
            assert cnt[f] == len(o["features"])


def test_config3_harris_3840x2160_tiled_and_untiled(ctx, checker):
    """configs[3]: Harris thr 30 on a 3840x2160 frame: candidate set bitwise vs the checker, 8 row tiles == untiled."""
    import torch
    from feature_detector_b200 import tiling
    im = synth(3840, 2160, 0)
    prm = fd.DetectParams(fd.HARRIS, 30.0, 20, 200)
    ctx.upload(im)
    ctx.detect(prm)
    kp_ref, cnt_ref = ctx.keypoints(200)
    cand = ctx.candidates(0)
    o = checker.detect(HARRIS, im, 30.0, 20, 200)
    assert len(cand) == o["n_cand"]
    want = np.sort(tiling.make_keys(o["cand_resp"], o["cand_xy"][:, 1], o["cand_xy"][:, 0]))
    assert np.array_equal(np.sort(tiling.make_keys(cand["response"], cand["y"], cand["x"])), want)
    kp, keys = tiling.detect_tiled_local(ctx, torch.from_numpy(im).cuda(), 8, prm)
    assert np.array_equal(np.sort(keys.cpu().numpy().view(np.uint64)), want)       # response bitwise across tile seams
    assert np.array_equal(kp, kp_ref[0, :cnt_ref[0]])
    if not np.array_equal(_feats(kp_ref, cnt_ref, 0), o["features"]):               # ties: same multiset, same count
        assert cnt_ref[0] == len(o["features"])


def test_config4_lsd_field_1920x1080(ctx, checker):
    """configs[4]: LSD norm bit-exact, angle within 1e-5 where valid, valid mask equal, seed order = norm descending."""
    frames = np.stack([synth(1920, 1080, 900 + i) for i in range(3)])
    ctx.upload(frames)
    ctx.lsd_field(fd.LsdParams(20.0, 1))
    for f in range(3):
        g = ctx.lsd_download(f)
        m = checker.lsd_map(frames[f])
        assert np.array_equal(g["norm"][:-1, :-1].view(np.uint32), m["norm"].view(np.uint32))
        valid = m["valid"].astype(bool)
        assert g["n_valid"] == int(valid.sum())
        assert np.max(np.abs(g["angle"][:-1, :-1][valid] - m["angle"][valid])) <= 1e-5
        assert not np.any(g["angle"][:-1, :-1][~valid])
        idx = g["sorted_idx"]
        r, c = idx // 1920, idx % 1920
        seq = g["norm"][r, c]
        assert np.all(np.diff(seq) <= 0)
        exp = m["sorted_rc"]                                                         # the reference's std::sort leaves ties open:
        assert np.array_equal(seq, m["norm"][exp[:, 0], exp[:, 1]])                  # same norm sequence,
        cm = c.astype(np.int64) * 1080 + r                                           # and ties here in its push order
        assert np.all(np.diff(cm)[np.diff(seq) == 0] > 0)
        assert len(np.unique(idx)) == len(idx)


def test_config3_second_frame_and_shi_tomasi_3840x2160(ctx, checker):
    """configs[3] again on other frames: Harris and Shi-Tomasi candidate sets bitwise at 3840x2160 (the frames bench.py times)."""
    from feature_detector_b200 import tiling
    for idx, kind, ck, thr in ((1, fd.HARRIS, HARRIS, 30.0), (2, fd.SHI_TOMAS, SHI_TOMAS, 40.0)):
        im = synth(3840, 2160, idx)
        ctx.upload(im)
        ctx.detect(fd.DetectParams(kind, thr, 20, 200))
        kp, cnt = ctx.keypoints(200)
        cand = ctx.candidates(0)
        o = checker.detect(ck, im, thr, 20, 200)
        assert len(cand) == o["n_cand"] and cnt[0] == len(o["features"])
        want = np.sort(tiling.make_keys(o["cand_resp"], o["cand_xy"][:, 1], o["cand_xy"][:, 0]))
        assert np.array_equal(np.sort(tiling.make_keys(cand["response"], cand["y"], cand["x"])), want)
