"""Fuzz of the CUDA path against the checker on random small frames -- the generator of
tests/test_oracle.py::test_port_equals_reference_on_random_small_inputs pointed at the C ABI: frames from 1 x 1 pixels up, extreme
thresholds, min distances 0 .. 100 (0 = cells of one pixel in the selection kernel), pre-existing features, LSD maps, BRIEF at
out-of-range keypoints."""
import numpy as np
import pytest

import feature_detector_b200 as fd

pytestmark = pytest.mark.gpu


def test_cuda_path_equals_checker_on_random_small_frames(checker):
    from oracle.bindings import FAST, HARRIS, SHI_TOMAS
    kinds = {FAST: fd.FAST, HARRIS: fd.HARRIS, SHI_TOMAS: fd.SHI_TOMAS}
    import os
    rng = np.random.default_rng(int(os.environ.get("FD_FUZZ_SEED", "20261018")))   # FD_FUZZ_CASES / FD_FUZZ_SEED: longer one-off runs
    with fd.Context(0) as ctx:
        for it in range(int(os.environ.get("FD_FUZZ_CASES", "400"))):
            rows, cols = int(rng.integers(1, 70)), int(rng.integers(1, 90))
            mode = int(rng.integers(0, 4))
            if mode == 0:
                img = rng.integers(0, 256, (rows, cols), dtype=np.uint8)
            elif mode == 1:
                img = (rng.integers(0, 2, (rows, cols)) * 255).astype(np.uint8)
            elif mode == 2:
                img = np.full((rows, cols), int(rng.integers(0, 256)), np.uint8)
            else:
                img = np.zeros((rows, cols), np.uint8)
                img[rows // 3:, cols // 4:] = 200
                img = (img + rng.integers(0, 6, (rows, cols))).astype(np.uint8)
            kind = (FAST, HARRIS, SHI_TOMAS)[int(rng.integers(0, 3))]
            thr = float(rng.choice([0.0, 0.1, 5.0, 10.0, 30.0, 1e9, -1.0]))
            d, needed, fast_n = int(rng.choice([0, 1, 5, 15, 20, 100])), int(rng.choice([0, 1, 3, 50, 1000])), int(rng.choice([9, 12]))
            n_pre = int(rng.choice([0, 0, 3]))
            pre = np.stack([rng.uniform(0, cols, n_pre), rng.uniform(0, rows, n_pre)], 1).astype(np.float32) if n_pre else None
            case = (it, rows, cols, kind, thr, d, needed, fast_n, n_pre)
            o = checker.detect(kind, img, thr, d, needed, fast_n=fast_n, pre=pre)
            ctx.upload(img)
            ctx.set_existing_features([pre] if n_pre else [])
            ctx.detect(fd.DetectParams(kinds[kind], thr, d, needed, fast_n=fast_n))
            kp, cnt = ctx.keypoints(max(needed, 1))
            cand = ctx.candidates(0)
            assert len(cand) == o["n_cand"], case
            g = np.lexsort((cand["x"], cand["y"]))
            c = np.lexsort((o["cand_xy"][:, 0], o["cand_xy"][:, 1]))
            assert np.array_equal(cand["x"][g], o["cand_xy"][c, 0]) and np.array_equal(cand["y"][g], o["cand_xy"][c, 1]), case
            assert np.array_equal(cand["response"][g].view(np.uint32), o["cand_resp"][c].view(np.uint32)), case
            got = np.stack([kp["x"][0, :cnt[0]], kp["y"][0, :cnt[0]]], 1).astype(np.float32)
            new = o["features"][n_pre:]
            if not np.array_equal(got, new):
                assert len(np.unique(o["cand_resp"])) < len(o["cand_resp"]), case      # only a tie may reorder the walk
            if it % 4 == 0:   # the one-call host-to-host form returns what the separate calls returned
                kp1, cnt1, _ = ctx.detect_describe_host(img, fd.DetectParams(kinds[kind], thr, d, needed, fast_n=fast_n), None, max(needed, 1))
                assert cnt1[0] == cnt[0] and np.array_equal(kp1[0, :cnt1[0]], kp[0, :cnt[0]]), case
            if rows >= 3 and cols >= 3:
                ctx.lsd_field(fd.LsdParams(20.0, 1))
                m = ctx.lsd_download(0)
                e = checker.lsd_map(img)
                assert np.array_equal(m["norm"][:-1, :-1].view(np.uint32), e["norm"].view(np.uint32)), case
                assert m["n_valid"] == int(e["valid"].sum()), case
                ang = np.where(e["valid"] != 0, e["angle"], 0.0)
                assert np.all(np.abs(m["angle"][:-1, :-1] - ang) <= 1e-5), case
            if rows > 40 and cols > 40:
                pts = np.stack([rng.uniform(-2, cols + 2, 6), rng.uniform(-2, rows + 2, 6)], 1).astype(np.float32)
                cap = ctx.describe_points(fd.BriefParams(256, 8), [pts])
                bits = fd.unpack_bits(ctx.descriptors(cap)[0, :len(pts)])
                assert np.array_equal(bits, checker.brief(img, pts, 256, 8)[1]), case
        ctx.set_existing_features([])


def test_cuda_path_equals_checker_on_random_medium_frames(checker):
    """The same comparison on frames of random size up to 1000 x 700 in batches of one to three: several column strips and row bands,
    ragged last strips, rank ranges in the selection (thousands of candidates), the per-cell form (FAST at a low threshold)."""
    import os
    from feature_detector_b200.synth import synth
    from oracle.bindings import FAST, HARRIS, SHI_TOMAS
    kinds = {FAST: fd.FAST, HARRIS: fd.HARRIS, SHI_TOMAS: fd.SHI_TOMAS}
    rng = np.random.default_rng(int(os.environ.get("FD_FUZZ_SEED", "20261018")) + 1)
    with fd.Context(0) as ctx:
        for it in range(int(os.environ.get("FD_FUZZ_MEDIUM_CASES", "60"))):
            rows, cols, n = int(rng.integers(40, 700)), int(rng.integers(40, 1000)), int(rng.integers(1, 4))
            if rng.integers(0, 3) == 0:
                frames = rng.integers(0, 256, (n, rows, cols), dtype=np.uint8)
            else:
                frames = np.stack([synth(cols, rows, int(rng.integers(0, 1000))) for _ in range(n)])
            kind = (FAST, HARRIS, SHI_TOMAS)[int(rng.integers(0, 3))]
            thr = float(rng.choice([0.1, 5.0, 10.0, 30.0, 40.0, 200.0]))
            d, needed, fast_n = int(rng.choice([1, 5, 15, 20, 60])), int(rng.choice([1, 50, 200, 1000])), int(rng.choice([9, 12]))
            case = (it, rows, cols, n, kind, thr, d, needed, fast_n)
            ctx.upload(frames)
            ctx.detect(fd.DetectParams(kinds[kind], thr, d, needed, fast_n=fast_n))
            kp, cnt = ctx.keypoints(needed)
            for f in range(n):
                o = checker.detect(kind, frames[f], thr, d, needed, fast_n=fast_n)
                cand = ctx.candidates(f)
                assert len(cand) == o["n_cand"], case
                g = np.lexsort((cand["x"], cand["y"]))
                c = np.lexsort((o["cand_xy"][:, 0], o["cand_xy"][:, 1]))
                assert np.array_equal(cand["x"][g], o["cand_xy"][c, 0]) and np.array_equal(cand["y"][g], o["cand_xy"][c, 1]), case
                assert np.array_equal(cand["response"][g].view(np.uint32), o["cand_resp"][c].view(np.uint32)), case
                got = np.stack([kp["x"][f, :cnt[f]], kp["y"][f, :cnt[f]]], 1).astype(np.float32)
                if not np.array_equal(got, o["features"]):
                    assert len(np.unique(o["cand_resp"])) < len(o["cand_resp"]), case      # only a tie may reorder the walk
                    assert len(got) == len(o["features"]) or needed > len(got), case
