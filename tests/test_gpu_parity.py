"""GPU suite: the CUDA path, called through the C ABI, against the CPU checker on the same inputs.

Bit-exact for FAST scores / responses, candidate sets, keypoint lists (raster tie rule), BRIEF bits, LSD
norms; Harris / Shi-Tomasi responses are compared bitwise too (BASELINE.json allows 1e-4 relative, SURVEY.md
H1 shows bitwise is reachable); LSD angles within 1e-5 (different atan2f implementations).
"""
import numpy as np
import pytest

import feature_detector_b200 as fd
from conftest import fnv
from oracle.bindings import FAST, HARRIS, SHI_TOMAS
from oracle.tiecheck import greedy_replay, same_up_to_ties

pytestmark = pytest.mark.gpu

KIND = {"fast": FAST, "fast9": FAST, "harris": HARRIS, "shi": SHI_TOMAS}


@pytest.fixture(scope="module")
def ctx():
    c = fd.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available()
    return torch


def _gpu_detect(ctx, im, kind, thr, d, n, fast_n=12, cap=0):
    ctx.upload(im)
    ctx.detect(fd.DetectParams(kind, thr, d, n, fast_n=fast_n), cap)
    kp, cnt = ctx.keypoints(max(n, 1))
    cand = ctx.candidates(0)
    feats = np.stack([kp["x"][0, :cnt[0]], kp["y"][0, :cnt[0]]], 1).astype(np.float32)
    return feats, cand, kp[0, :cnt[0]]


def _assert_same_candidates(cand, o):
    """GPU candidates (sorted: response desc, raster ties) vs checker candidates (its own sorted order)."""
    assert len(cand) == o["n_cand"]
    g = np.lexsort((cand["x"], cand["y"]))
    c = np.lexsort((o["cand_xy"][:, 0], o["cand_xy"][:, 1]))
    assert np.array_equal(cand["x"][g], o["cand_xy"][c, 0]) and np.array_equal(cand["y"][g], o["cand_xy"][c, 1])
    assert np.array_equal(cand["response"][g].view(np.uint32), o["cand_resp"][c].view(np.uint32))
    # and the GPU's own order is (response desc, y, x)
    order = np.lexsort((cand["x"], cand["y"], -cand["response"].astype(np.float64)))
    assert np.array_equal(order, np.arange(len(cand)))


def _assert_same_features(feats, cand, o, shape, d, n):
    if np.array_equal(feats, o["features"]):
        return
    # identical except for ties: same sorted multiset, and each list is the greedy replay of its own order
    assert same_up_to_ties(o["cand_resp"], o["cand_xy"], cand["response"], np.stack([cand["x"], cand["y"]], 1))
    replay, _ = greedy_replay(np.stack([cand["x"], cand["y"]], 1), shape[0], shape[1], d, n)
    assert np.array_equal(replay, feats)


def test_golden_detector_cases(ctx, frames, kat):
    """Config 0 of BASELINE.json (image.png, the demo option values) and the synthetic goldens, no oracle needed."""
    n = 0
    for c in kat["cases"]:
        if c["detector"] not in KIND:
            continue
        im = frames[c["frame"]]
        feats, cand, _ = _gpu_detect(ctx, im, KIND[c["detector"]], c["thr"], c["dist"], c["needed"], c["fast_n"])
        assert len(cand) == c["n_cand"], c
        g = np.lexsort((cand["x"], cand["y"]))
        flat = np.zeros((len(cand), 3), np.uint32)
        flat[:, 0] = cand["response"][g].view(np.uint32)
        flat[:, 1] = cand["x"][g]
        flat[:, 2] = cand["y"][g]
        assert fnv(flat) == c["cand_hash"], c
        assert len(feats) == c["n_feat_raster_ties"]
        assert fnv(feats.astype("<i4")) == c["feat_hash_raster_ties"], c
        n += 1
    assert n >= 21


@pytest.mark.parametrize("shape_idx", [(752, 480, 21), (333, 217, 5), (1280, 720, 2), (130, 70, 9), (64, 48, 1), (1001, 37, 3)])
@pytest.mark.parametrize("case", [("fast", 10.0, 20, 200, 12), ("fast", 10.0, 20, 200, 9), ("fast", 0.1, 15, 300, 12), ("fast", 3.0, 7, 100, 9),
                                  ("harris", 30.0, 20, 200, 12), ("harris", 0.1, 15, 1000, 12), ("shi", 40.0, 20, 1000, 12),
                                  ("shi", 0.1, 4, 3000, 12)])
def test_detect_vs_checker(ctx, checker, shape_idx, case):
    from feature_detector_b200.synth import synth
    w, h, idx = shape_idx
    name, thr, d, n, fast_n = case
    im = synth(w, h, idx)
    o = checker.detect(KIND[name], im, thr, d, n, fast_n=fast_n)
    feats, cand, _ = _gpu_detect(ctx, im, KIND[name], thr, d, n, fast_n)
    _assert_same_candidates(cand, o)
    _assert_same_features(feats, cand, o, im.shape, d, n)


def test_dense_maps_vs_checker(ctx, checker, torch_cuda, frames):
    torch = torch_cuda
    for name, im in frames.items():
        h, w = im.shape
        resp = torch.full((h, w), -7.0, dtype=torch.float32, device="cuda")
        score = torch.full((h, w), 99, dtype=torch.uint8, device="cuda")
        ctx.upload(im)
        ctx.set_dense_outputs(resp.data_ptr(), score.data_ptr())
        try:
            for kind, thr in ((HARRIS, 0.1), (HARRIS, 30.0), (SHI_TOMAS, 0.1), (SHI_TOMAS, 40.0)):
                ctx.compute_candidates(fd.DetectParams(kind, thr, 15, 10))
                ctx.sync()
                o = checker.detect(kind, im, thr, 15, 10, want_response=True, want_candidates=False)
                got = resp.cpu().numpy()
                assert np.array_equal(got.view(np.uint32), o["response"].view(np.uint32)), (name, kind, thr)
            for fast_n in (12, 9):
                ctx.compute_candidates(fd.DetectParams(FAST, 10.0, 15, 10, fast_n=fast_n))
                ctx.sync()
                assert np.array_equal(score.cpu().numpy(), checker.fast_score_map(im, fast_n, 15)), (name, fast_n)
        finally:
            ctx.set_dense_outputs(0, 0)


def test_fast_pixel_diff_parameter(ctx, checker, torch_cuda):
    from feature_detector_b200.synth import synth
    torch = torch_cuda
    im = synth(320, 200, 8)
    score = torch.zeros(im.shape, dtype=torch.uint8, device="cuda")
    ctx.upload(im)
    ctx.set_dense_outputs(0, score.data_ptr())
    try:
        for diff in (0, 1, 15, 40, 255):
            ctx.compute_candidates(fd.DetectParams(FAST, 10.0, 15, 10, fast_n=9, fast_min_pixel_diff=diff))
            ctx.sync()
            assert np.array_equal(score.cpu().numpy(), checker.fast_score_map(im, 9, diff)), diff
    finally:
        ctx.set_dense_outputs(0, 0)


def test_batch_equals_single_frames(ctx, checker):
    """A batch is processed frame by frame identically (frames are independent)."""
    from feature_detector_b200.synth import synth_batch
    batch = synth_batch(752, 480, 6, start=100)
    for kind, thr, d, n, fast_n in ((FAST, 10.0, 20, 200, 9), (HARRIS, 30.0, 20, 200, 12)):
        ctx.upload(batch)
        ctx.detect(fd.DetectParams(kind, thr, d, n, fast_n=fast_n))
        kp, cnt = ctx.keypoints(n)
        counts = ctx.candidate_counts()
        for f in range(len(batch)):
            o = checker.detect(kind, batch[f], thr, d, n, fast_n=fast_n)
            assert counts[f] == o["n_cand"]
            feats = np.stack([kp["x"][f, :cnt[f]], kp["y"][f, :cnt[f]]], 1)
            cand = ctx.candidates(f)
            _assert_same_candidates(cand, o)
            _assert_same_features(feats, cand, o, batch[f].shape, d, n)


def test_candidate_capacity_overflow_is_reported(ctx):
    from feature_detector_b200.synth import synth
    ctx.upload(synth(752, 480, 0))
    ctx.detect(fd.DetectParams(HARRIS, 30.0, 20, 200), 1000)  # ~23k candidates on this frame
    with pytest.raises(fd.FdError) as e:
        ctx.keypoints(200)
    assert e.value.status == 5


def test_needed_zero_and_flat_images(ctx, checker):
    from feature_detector_b200.synth import synth
    im = synth(160, 120, 3)
    feats, _, _ = _gpu_detect(ctx, im, HARRIS, 30.0, 20, 0)
    assert len(feats) == 1 and np.array_equal(feats, checker.detect(HARRIS, im, 30.0, 20, 0)["features"])
    flat = np.full((64, 64), 77, np.uint8)
    for kind, thr in ((HARRIS, 0.1), (SHI_TOMAS, 0.1), (FAST, 10.0)):
        feats, cand, _ = _gpu_detect(ctx, flat, kind, thr, 15, 100)
        assert len(feats) == 0 and len(cand) == 0
    feats, cand, _ = _gpu_detect(ctx, flat, FAST, 0.001, 15, 100)
    o = checker.detect(FAST, flat, 0.001, 15, 100)
    _assert_same_candidates(cand, o)
    assert np.array_equal(feats, o["features"])
    for shape in [(4, 4), (6, 9), (7, 7), (5, 5)]:
        tiny = np.arange(shape[0] * shape[1], dtype=np.uint8).reshape(shape) * 9
        for kind in (HARRIS, SHI_TOMAS, FAST):
            feats, cand, _ = _gpu_detect(ctx, tiny, kind, 1.0, 2, 10)
            o = checker.detect(kind, tiny, 1.0, 2, 10)
            _assert_same_candidates(cand, o)


def test_brief_golden_and_checker(ctx, checker, frames, kat, vectors):
    for c in kat["cases"]:
        if c["detector"] != "brief":
            continue
        im = frames[c["frame"]]
        kp = vectors[f"{c['frame']}.brief.{c['set']}.kp"]
        ctx.upload(im)
        cap = ctx.describe_points(fd.BriefParams(c["length"], 8), [kp])
        desc = ctx.descriptors(cap)[0, :len(kp)]
        assert np.array_equal(desc[:, :(c["length"] + 7) // 8], vectors[f"{c['frame']}.brief.{c['set']}.desc"]), c
        assert not desc[:, (c["length"] + 7) // 8:].any()
        assert fnv(np.ascontiguousarray(desc[:, :(c["length"] + 7) // 8])) == c["hash"]
    # other patch sizes / lengths and keypoints right at the border limit, against the checker
    from feature_detector_b200.synth import synth
    im = synth(400, 300, 12)
    rng = np.random.default_rng(2)
    kp = np.stack([rng.uniform(0, 400, 200), rng.uniform(0, 280, 200)], 1).astype(np.float32)
    kp[:40] = np.floor(kp[:40])
    kp[40] = (19.0, 19.0)
    kp[41] = (18.999, 50.0)
    kp[42] = (400 - 19.0, 100.0)
    for length, hp in ((256, 8), (200, 12), (31, 3), (256, 0)):
        ctx.upload(im)
        cap = ctx.describe_points(fd.BriefParams(length, hp), [kp])
        bits = fd.unpack_bits(ctx.descriptors(cap)[0, :len(kp)], length)
        ok, exp = checker.brief(im, kp, length, hp)
        assert ok and np.array_equal(bits, exp), (length, hp)


def test_brief_float_overload(ctx, checker, frames):
    """The std::vector<Vec> overload of Descriptor::Compute (descriptor.h:43-62) on the device: +1 / -1 per bit."""
    for name, im in frames.items():
        for length in (256, 100):
            ctx.upload(im)
            ctx.detect(fd.DetectParams(fd.HARRIS, 20.0, 20, 80))
            ctx.describe_selected(fd.BriefParams(length, 8))
            kp, cnt = ctx.keypoints(80)
            got = ctx.descriptors_float(80, length)[0, :cnt[0]]
            pts = np.stack([kp["x"][0, :cnt[0]], kp["y"][0, :cnt[0]]], 1).astype(np.float32)
            ok, exp = checker.brief_vec(im, pts, length)
            assert ok and np.array_equal(got, exp), (name, length)


def test_detect_then_describe_stays_on_device(ctx, checker, image_png):
    """FAST -> select -> BRIEF without a host round trip (config 1 pipeline), against the checker run stage by stage."""
    ctx.upload(image_png)
    ctx.detect(fd.DetectParams(FAST, 10.0, 20, 200, fast_n=12))
    ctx.describe_selected(fd.BriefParams(256, 8))
    kp, cnt = ctx.keypoints(200)
    desc = ctx.descriptors(200)
    o = checker.detect(FAST, image_png, 10.0, 20, 200)
    feats = np.stack([kp["x"][0, :cnt[0]], kp["y"][0, :cnt[0]]], 1)
    assert np.array_equal(feats, o["features"]) and cnt[0] == 84
    ok, exp = checker.brief(image_png, o["features"], 256, 8)
    assert np.array_equal(fd.unpack_bits(desc[0, :cnt[0]]), exp)


def test_lsd_field_vs_checker(ctx, checker, frames):
    from feature_detector_b200.synth import synth
    cases = dict(frames)
    cases["s1920"] = synth(1920, 1080, 1)
    cases["tiny"] = synth(9, 7, 1)
    for name, im in cases.items():
        h, w = im.shape
        ctx.upload(im)
        ctx.lsd_field(fd.LsdParams(20.0, 1))
        g = ctx.lsd_download(0)
        o = checker.lsd_map(im)
        norm, angle = g["norm"][:h - 1, :w - 1], g["angle"][:h - 1, :w - 1]
        assert not g["norm"][h - 1:].any() and not g["norm"][:, w - 1:].any()
        assert np.array_equal(norm.view(np.uint32), o["norm"].view(np.uint32)), name           # bit-exact
        valid = norm > 20.0
        assert np.array_equal(valid, o["valid"].astype(bool))
        assert not angle[~valid].any()
        assert np.max(np.abs(angle[valid] - o["angle"][valid]), initial=0.0) <= 1e-5, name       # north_star tolerance
        assert g["n_valid"] == int(o["valid"].sum())
        # seed order: norm descending; ties in column-major push order (= a stable sort of the reference's push order)
        rows_, cols_ = g["sorted_idx"] // w, g["sorted_idx"] % w
        seq = norm[rows_, cols_]
        assert np.all(np.diff(seq) <= 0)
        exp_rc = o["sorted_rc"]
        assert np.array_equal(seq, o["norm"][exp_rc[:, 0], exp_rc[:, 1]])
        cm = cols_.astype(np.int64) * h + rows_
        same = np.diff(seq) == 0
        assert np.all(np.diff(cm)[same] > 0)
        assert len(np.unique(g["sorted_idx"])) == len(g["sorted_idx"])


def _gradient_pair_image(sign_ad, sign_bc):
    """Every (ad, bc) pair of feature_line_detector.cpp:76-79 with the given signs on 8-bit pixels: at even columns of even
    rows the 2x2 neighbourhood [[a, b], [c, d]] has |d - a| = D = row / 2 and |b - c| = B = column / 2."""
    im = np.zeros((514, 516), np.uint8)
    b = (np.arange(516) // 2).clip(0, 255).astype(np.uint8)[None, :] * np.ones((514, 1), np.uint8)
    d = (np.arange(514) // 2).clip(0, 255).astype(np.uint8)[:, None] * np.ones((1, 516), np.uint8)
    if sign_ad > 0:
        im[1::2, 1::2] = d[1::2, 1::2]     # d = D, a = 0
    else:
        im[0::2, 0::2] = d[0::2, 0::2]     # a = D, d = 0
    if sign_bc > 0:
        im[0::2, 1::2] = b[0::2, 1::2]     # b = B, c = 0
    else:
        im[1::2, 0::2] = b[1::2, 0::2]     # c = B, b = 0
    return im


def test_lsd_norm_every_gradient_pair(ctx, checker):
    """The field kernel's own square root (reciprocal-sqrt seed + one fused residual step) must round exactly like sqrtf,
    and its own arctangent stay inside the angle budget, for every attainable (ad, bc); odd widths take the scalar store
    path, min_norm 0 makes nearly every pixel valid (full queues)."""
    seen = set()
    worst = 0.0
    for k, (sa, sb) in enumerate(((1, 1), (-1, 1), (1, -1), (-1, -1))):
        im = _gradient_pair_image(sa, sb)
        ad = im[1:, 1:].astype(np.int32) - im[:-1, :-1]
        bc = im[:-1, 1:].astype(np.int32) - im[1:, :-1]
        seen |= set(zip(ad[1:-1, 1:-1].ravel().tolist(), bc[1:-1, 1:-1].ravel().tolist()))
        img, thr = ((im, 20.0), (im[:, :515], 0.0), (im.T, 300.0), (im[:513], 1.0))[k]
        img = np.ascontiguousarray(img)
        h, w = img.shape
        ctx.upload(img)
        ctx.lsd_field(fd.LsdParams(thr, 1))
        g = ctx.lsd_download(0)
        o = checker.lsd_map(img, thr)
        norm, angle = g["norm"][:h - 1, :w - 1], g["angle"][:h - 1, :w - 1]
        assert np.array_equal(norm.view(np.uint32), o["norm"].view(np.uint32))
        valid = o["valid"].astype(bool)
        assert g["n_valid"] == int(valid.sum())
        assert not angle[~valid].any()
        worst = max(worst, float(np.max(np.abs(angle[valid].astype(np.float64) - o["angle"][valid]), initial=0.0)))
        rows_, cols_ = g["sorted_idx"] // w, g["sorted_idx"] % w
        assert np.array_equal(norm[rows_, cols_], o["norm"][o["sorted_rc"][:, 0], o["sorted_rc"][:, 1]])
    assert len(seen) >= 511 * 511
    assert worst <= 1e-6, worst   # budget 1e-5 (BASELINE.json north_star); the kernel's arctangent is good to 5e-7


def test_lsd_seed_order_of_a_ramp(ctx, checker):
    """A linear ramp gives every pixel the same gradient: one magnitude bucket holds (nearly) all seeds, which takes the
    in-place sort of the ordering kernel instead of its shared-memory tiles."""
    r, c = np.mgrid[0:300, 0:420]
    img = ((5 * r + 3 * c) & 0xFF).astype(np.uint8)
    h, w = img.shape
    ctx.upload(img)
    ctx.lsd_field(fd.LsdParams(2.0, 1))
    g = ctx.lsd_download(0)
    o = checker.lsd_map(img, 2.0)
    norm = g["norm"][:h - 1, :w - 1]
    assert np.array_equal(norm.view(np.uint32), o["norm"].view(np.uint32))
    assert g["n_valid"] == int(o["valid"].sum()) > 100000
    rows_, cols_ = g["sorted_idx"] // w, g["sorted_idx"] % w
    seq = norm[rows_, cols_]
    assert np.array_equal(seq, o["norm"][o["sorted_rc"][:, 0], o["sorted_rc"][:, 1]])
    cm = cols_.astype(np.int64) * h + rows_
    assert np.all(np.diff(cm)[np.diff(seq) == 0] > 0)          # ties in the reference's push order: column outer, row inner
    assert len(np.unique(g["sorted_idx"])) == len(g["sorted_idx"])


def test_bound_device_frames_with_pitch(ctx, checker, torch_cuda):
    """Frames that already live on the device, including an unaligned pitch (re-pitched internally)."""
    from feature_detector_b200.synth import synth_batch
    torch = torch_cuda
    batch = synth_batch(333, 217, 3, start=40)
    for pitch in (333, 336, 400):
        buf = torch.zeros((3, 217, pitch), dtype=torch.uint8, device="cuda")
        buf[:, :, :333] = torch.from_numpy(batch).cuda()
        ctx.bind_device(buf.data_ptr(), 217, 333, 3, pitch=pitch)
        ctx.detect(fd.DetectParams(FAST, 10.0, 20, 200, fast_n=9))
        kp, cnt = ctx.keypoints(200)
        for f in range(3):
            o = checker.detect(FAST, batch[f], 10.0, 20, 200, fast_n=9)
            assert np.array_equal(np.stack([kp["x"][f, :cnt[f]], kp["y"][f, :cnt[f]]], 1), o["features"])


# ---- pre-existing features: the in/out `features` argument (feature_point_detector.cpp:12-16, 90-98) -------------
def _pre_sets(w, h, rng):
    grid = np.array([[15.0 * i, 15.0 * j] for i in range(1, 10) for j in range(1, 10)], np.float32)  # the demo's 81 seeds
    rand = np.stack([rng.uniform(-30, w + 30, 40), rng.uniform(-30, h + 30, 40)], 1).astype(np.float32)  # fractional, some outside
    return {"grid81": grid, "random_fractional": rand, "corner_cases": np.array([[0, 0], [w - 1, h - 1], [w + 5, 3], [-0.5, -0.5], [3.9, 3.9]], np.float32)}


@pytest.mark.parametrize("shape_idx", [(752, 480, 0), (333, 217, 5), (130, 70, 9)])
@pytest.mark.parametrize("case", [("fast", 10.0, 20, 200, 12), ("fast", 10.0, 20, 200, 9), ("fast", 0.1, 15, 300, 12), ("fast", 0.5, 6, 500, 9),
                                  ("harris", 30.0, 20, 200, 12), ("harris", 0.1, 15, 1000, 12), ("shi", 40.0, 20, 300, 12)])
def test_detect_with_existing_features_vs_checker(ctx, checker, shape_idx, case):
    from feature_detector_b200.synth import synth
    w, h, idx = shape_idx
    name, thr, d, n, fast_n = case
    im = synth(w, h, idx)
    rng = np.random.default_rng(idx)
    try:
        for label, pre in _pre_sets(w, h, rng).items():
            o = checker.detect(KIND[name], im, thr, d, n, fast_n=fast_n, pre=pre)
            ctx.upload(im)
            ctx.set_existing_features([pre])
            ctx.detect(fd.DetectParams(KIND[name], thr, d, n, fast_n=fast_n))
            kp, cnt = ctx.keypoints(max(n, 1))
            cand = ctx.candidates(0)
            new = np.stack([kp["x"][0, :cnt[0]], kp["y"][0, :cnt[0]]], 1).astype(np.float32)
            _assert_same_candidates(cand, o)
            assert np.array_equal(o["features"][:len(pre)], pre), label       # the reference keeps the seeds in front
            ref_new = o["features"][len(pre):]
            if not np.array_equal(new, ref_new):
                assert name != "fast", label                                   # FAST has no ties
                assert same_up_to_ties(o["cand_resp"], o["cand_xy"], cand["response"], np.stack([cand["x"], cand["y"]], 1)), label
                assert len(new) == len(ref_new), label
    finally:
        ctx.set_existing_features([])


def test_existing_features_count_toward_needed(ctx, checker, image_png):
    """needed <= len(existing): the reference still pushes exactly one new feature (push, then test; :67-68)."""
    pre = np.array([[15.0 * i, 15.0 * j] for i in range(1, 10) for j in range(1, 10)], np.float32)
    try:
        for needed in (0, 50, 81, 82, 100):
            o = checker.detect(HARRIS, image_png, 30.0, 20, needed, pre=pre)
            ctx.upload(image_png)
            ctx.set_existing_features([pre])
            ctx.detect(fd.DetectParams(fd.HARRIS, 30.0, 20, needed))
            kp, cnt = ctx.keypoints(max(needed, 1))
            new = np.stack([kp["x"][0, :cnt[0]], kp["y"][0, :cnt[0]]], 1).astype(np.float32)
            assert np.array_equal(new, o["features"][len(pre):]), needed
    finally:
        ctx.set_existing_features([])


def test_golden_preseeded_harris(ctx, image_png, kat):
    """Harris thr 30 + 81 pre-seeded (15i, 15j): 12 194 candidates, 119 new features, first new (520, 201)."""
    pre = np.array([[15.0 * i, 15.0 * j] for i in range(1, 10) for j in range(1, 10)], np.float32)
    try:
        ctx.upload(image_png)
        ctx.set_existing_features([pre])
        ctx.detect(fd.DetectParams(fd.HARRIS, 30.0, 20, 200))
        kp, cnt = ctx.keypoints(200)
        assert int(ctx.candidate_counts()[0]) == 12194
        assert cnt[0] == 119 and (kp["x"][0, 0], kp["y"][0, 0]) == (520.0, 201.0)
        full = np.concatenate([pre, np.stack([kp["x"][0, :cnt[0]], kp["y"][0, :cnt[0]]], 1)]).astype("<i4")
        assert fnv(full) == "85bd4f625b5448b3"
    finally:
        ctx.set_existing_features([])


def test_existing_features_per_frame_in_a_batch(ctx, checker):
    from feature_detector_b200.synth import synth
    ims = np.stack([synth(320, 200, i) for i in range(4)])
    rng = np.random.default_rng(7)
    pres = [np.stack([rng.uniform(0, 320, k), rng.uniform(0, 200, k)], 1).astype(np.float32) for k in (0, 5, 60, 1)]
    try:
        ctx.upload(ims)
        ctx.set_existing_features(pres)
        ctx.detect(fd.DetectParams(fd.FAST, 0.3, 10, 150, fast_n=9))
        kp, cnt = ctx.keypoints(150)
        for f in range(4):
            o = checker.detect(FAST, ims[f], 0.3, 10, 150, fast_n=9, pre=pres[f])
            new = np.stack([kp["x"][f, :cnt[f]], kp["y"][f, :cnt[f]]], 1).astype(np.float32)
            assert np.array_equal(new, o["features"][len(pres[f]):]), f
    finally:
        ctx.set_existing_features([])


# ---- sparse FAST kernel (fd_fast_sparse.cu): same candidates as the dense kernel / the reference, bit for bit ----------
@pytest.fixture(scope="module")
def dense_ctx():
    """A context pinned to the dense FAST kernel (FD_B200_FAST_DENSE is read when the context is created)."""
    import os
    os.environ["FD_B200_FAST_DENSE"] = "1"
    try:
        c = fd.Context(0)
    finally:
        del os.environ["FD_B200_FAST_DENSE"]
    yield c
    c.close()


def _cand_table(ctx, frame=0):
    c = ctx.candidates(frame)
    t = np.zeros((len(c), 3), np.uint32)
    t[:, 0] = c["response"].view(np.uint32)
    t[:, 1] = c["x"]
    t[:, 2] = c["y"]
    return t


@pytest.mark.parametrize("shape_idx", [(752, 480, 3), (333, 217, 5), (130, 70, 9), (1001, 37, 3), (64, 48, 1), (1280, 720, 4), (2048, 131, 6)])
@pytest.mark.parametrize("fast_n", [9, 12])
def test_fast_sparse_thresholds_vs_checker(ctx, checker, shape_idx, fast_n):
    """Thresholds that put s_min anywhere in 1..17 (the sparse kernel's ANY / ADJ / PRE word tests and the switch rows)."""
    from feature_detector_b200.synth import synth
    w, h, idx = shape_idx
    im = synth(w, h, idx)
    ctx.upload(im)
    for thr in (4.05, 5.0, 7.0, 7.99, 8.0, 9.5, 10.0, 12.0, 15.9, 16.5, 1.0, 0.9):
        o = checker.detect(FAST, im, thr, 20, 100, fast_n=fast_n)
        ctx.detect(fd.DetectParams(fd.FAST, thr, 20, 100, fast_n=fast_n))
        _assert_same_candidates(ctx.candidates(0), o)
        kp, cnt = ctx.keypoints(100)
        assert np.array_equal(np.stack([kp["x"][0, :cnt[0]], kp["y"][0, :cnt[0]]], 1), o["features"]), (thr, fast_n)


@pytest.mark.parametrize("fast_n", [9, 12])
def test_fast_sparse_equals_dense_other_diffs_and_batches(ctx, dense_ctx, fast_n):
    from feature_detector_b200.synth import synth
    rng = np.random.default_rng(5)
    batches = [np.stack([synth(320, 200, i) for i in range(24)]),
               rng.integers(0, 256, (5, 97, 259), dtype=np.uint8),                      # pure noise: nearly every word survives
               np.stack([synth(752, 480, 40 + i) for i in range(3)])]
    for frames in batches:
        ctx.upload(frames)
        dense_ctx.upload(frames)
        for diff, thr in ((1, 9.0), (7, 6.0), (15, 10.0), (16, 10.0), (31, 5.0), (40, 8.5), (100, 4.5), (255, 4.5), (0, 12.0)):
            prm = fd.DetectParams(fd.FAST, thr, 12, 50, fast_n=fast_n, fast_min_pixel_diff=diff)
            ctx.detect(prm)
            dense_ctx.detect(prm)
            assert np.array_equal(ctx.candidate_counts(), dense_ctx.candidate_counts()), (diff, thr)
            for f in range(len(frames)):
                assert np.array_equal(_cand_table(ctx, f), _cand_table(dense_ctx, f)), (diff, thr, f)
            kp_a, cnt_a = ctx.keypoints(50)
            kp_b, cnt_b = dense_ctx.keypoints(50)
            assert np.array_equal(cnt_a, cnt_b) and np.array_equal(kp_a, kp_b)
        # with pre-existing features the sparse form takes the pixel index from the mask's prefix counts
        h, w = frames.shape[1:]
        existing = [np.stack([rng.integers(0, w, 12 + f), rng.integers(0, h, 12 + f)], 1).astype(np.float32) for f in range(len(frames))]
        ctx.set_existing_features(existing)
        dense_ctx.set_existing_features(existing)
        for diff, thr, d in ((15, 10.0, 20), (15, 8.5, 5), (40, 6.0, 31)):
            prm = fd.DetectParams(fd.FAST, thr, d, 60, fast_n=fast_n, fast_min_pixel_diff=diff)
            ctx.detect(prm)
            dense_ctx.detect(prm)
            assert np.array_equal(ctx.candidate_counts(), dense_ctx.candidate_counts()), (diff, thr, d)
            for f in range(len(frames)):
                assert np.array_equal(_cand_table(ctx, f), _cand_table(dense_ctx, f)), (diff, thr, d, f)
            kp_a, cnt_a = ctx.keypoints(60)
            kp_b, cnt_b = dense_ctx.keypoints(60)
            assert np.array_equal(cnt_a, cnt_b)
            for f in range(len(frames)):
                assert np.array_equal(kp_a[f, :cnt_a[f]], kp_b[f, :cnt_b[f]]), (diff, thr, d, f)
        ctx.set_existing_features([])
        dense_ctx.set_existing_features([])


# ---- the two corner kernels (fd_corner_tma.cu: TMA row ring; fd_corner.cu: register streaming) --------------------------------
def test_corner_kernels_agree_with_and_without_masks(ctx):
    """The TMA form is the default; the streaming form serves frames a tensor map cannot describe.  Same candidates and
    keypoints from both, with and without pre-existing features (the response is 0 where the mask is clear, harris.cpp:94)."""
    import os
    from feature_detector_b200.synth import synth
    os.environ["FD_B200_CORNER_STREAM"] = "1"
    try:
        stream_ctx = fd.Context(0)
    finally:
        del os.environ["FD_B200_CORNER_STREAM"]
    try:
        rng = np.random.default_rng(17)
        for frames in (np.stack([synth(320, 200, 30 + i) for i in range(5)]), np.stack([synth(752, 480, 90 + i) for i in range(2)]),
                       rng.integers(0, 256, (3, 67, 131), dtype=np.uint8)):
            h, w = frames.shape[1:]
            existing = [np.stack([rng.integers(0, w, 10 + f), rng.integers(0, h, 10 + f)], 1).astype(np.float32) for f in range(len(frames))]
            for with_existing in (False, True):
                for kind, thr, d in ((fd.HARRIS, 30.0, 20), (fd.SHI_TOMAS, 40.0, 9), (fd.HARRIS, 0.1, 15)):
                    tables, kps = [], []
                    for c in (ctx, stream_ctx):
                        c.upload(frames)
                        c.set_existing_features(existing if with_existing else [])
                        c.detect(fd.DetectParams(kind, thr, d, 150), 0)
                        tables.append([_cand_table(c, f) for f in range(len(frames))])
                        kps.append(c.keypoints(150))
                    for f in range(len(frames)):
                        assert np.array_equal(tables[0][f], tables[1][f]), (kind, thr, with_existing, f)
                        n = kps[0][1][f]
                        assert n == kps[1][1][f] and np.array_equal(kps[0][0][f, :n], kps[1][0][f, :n]), (kind, thr, with_existing, f)
        ctx.set_existing_features([])
    finally:
        stream_ctx.close()


# ---- the two forms of the selection rounds (fd_select.cu): per live candidate and per cell -------------------------------
def _ctx_with_select_threshold(value):
    import os
    os.environ["FD_B200_SELECT_CELLS_MIN"] = str(value)
    try:
        return fd.Context(0)
    finally:
        del os.environ["FD_B200_SELECT_CELLS_MIN"]


def test_selection_per_cell_equals_per_candidate(checker):
    """Frames below the switch-over count are selected per candidate, frames above it per cell; forcing either form on the
    same candidates must give the same keypoints, with and without pre-existing features, for one and for several rank
    batches, fine and coarse cell grids."""
    from feature_detector_b200.synth import synth
    per_cell, per_cand, default = _ctx_with_select_threshold(0), _ctx_with_select_threshold(1 << 30), fd.Context(0)
    try:
        rng = np.random.default_rng(11)
        batches = [np.stack([synth(320, 200, i) for i in range(6)]),
                   rng.integers(0, 256, (3, 120, 200), dtype=np.uint8),
                   np.stack([synth(752, 480, 70 + i) for i in range(2)])]
        cases = [(fd.HARRIS, 0.1, 15, 200, 12), (fd.HARRIS, 30.0, 3, 5000, 12), (fd.SHI_TOMAS, 0.1, 40, 1000, 12), (fd.FAST, 0.1, 15, 200, 12),
                 (fd.FAST, 10.0, 20, 200, 9), (fd.FAST, 0.1, 0, 300, 9), (fd.HARRIS, 0.1, 1, 100000, 12), (fd.FAST, 0.1, 7, 0, 12)]
        for frames in batches:
            h, w = frames.shape[1:]
            existing = [np.stack([rng.integers(0, w, 9), rng.integers(0, h, 9)], 1).astype(np.float32) for _ in range(len(frames))]
            for kind, thr, d, n, fast_n in cases:
                for with_existing in (False, True):
                    out = []
                    for c in (per_cell, per_cand, default):
                        c.upload(frames)
                        if with_existing:
                            c.set_existing_features(existing)
                        else:
                            c.set_existing_features([])
                        c.detect(fd.DetectParams(kind, thr, d, n, fast_n=fast_n), 0)
                        kp, cnt = c.keypoints(max(n, 1))
                        out.append((kp, cnt))
                    for other in (1, 2):
                        assert np.array_equal(out[0][1], out[other][1]), (kind, thr, d, n, with_existing, other)
                        for f in range(len(frames)):
                            k = out[0][1][f]
                            assert np.array_equal(out[0][0][f, :k], out[other][0][f, :k]), (kind, thr, d, n, with_existing, f, other)
            # and against the checker for one case per batch (ties are open in the reference's unstable sort)
            per_cell.upload(frames)
            per_cell.set_existing_features([])
            per_cell.detect(fd.DetectParams(fd.FAST, 0.1, 15, 200, fast_n=12), 0)
            kp, cnt = per_cell.keypoints(200)
            for f in range(len(frames)):
                o = checker.detect(FAST, frames[f], 0.1, 15, 200, fast_n=12)
                feats = np.stack([kp["x"][f, :cnt[f]], kp["y"][f, :cnt[f]]], 1).astype(np.float32)
                assert np.array_equal(feats, o["features"]), f
            # min distance 0 and 1: cells of one and two pixels (the cell index is the pixel itself at 0), both forms against the checker
            for d in (0, 1):
                for c in (per_cell, per_cand, default):
                    c.upload(frames)
                    c.set_existing_features([])
                    c.detect(fd.DetectParams(fd.FAST, 5.0, d, 400, fast_n=9), 0)
                    kp, cnt = c.keypoints(400)
                    for f in range(len(frames)):
                        o = checker.detect(FAST, frames[f], 5.0, d, 400, fast_n=9)
                        feats = np.stack([kp["x"][f, :cnt[f]], kp["y"][f, :cnt[f]]], 1).astype(np.float32)
                        assert np.array_equal(feats, o["features"]), (d, f)
    finally:
        per_cell.close()
        per_cand.close()
        default.close()


def test_host_pipeline_equals_single_call(ctx, torch_cuda):
    """Chunked, double-buffered host batches (pipeline.HostPipeline) give what one fd_detect over the batch gives."""
    from feature_detector_b200.pipeline import HostPipeline
    from feature_detector_b200.synth import synth
    torch = torch_cuda
    frames = np.stack([synth(320, 200, 100 + i) for i in range(21)])
    host = torch.from_numpy(frames).pin_memory()
    det, brief = fd.DetectParams(fd.FAST, 10.0, 12, 60, fast_n=9), fd.BriefParams(256, 8)
    ctx.upload(frames)
    ctx.detect(det)
    ctx.describe_selected(brief)
    kp_ref, cnt_ref = ctx.keypoints(60)
    desc_ref = ctx.descriptors(60)
    kp = np.zeros((21, 60), fd.KEYPOINT_DTYPE)
    cnt = np.zeros(21, np.int32)
    desc = np.zeros((21, 60, 32), np.uint8)
    with HostPipeline(0, chunk_frames=4) as pipe:   # 6 chunks, the last one ragged
        for _ in range(2):                            # second pass reuses the contexts' buffers
            pipe.run(host.data_ptr(), 200, 320, 21, det, brief, kp, cnt, desc)
            assert np.array_equal(cnt, cnt_ref)
            for f in range(21):
                assert np.array_equal(kp[f, :cnt[f]], kp_ref[f, :cnt[f]]) and np.array_equal(desc[f, :cnt[f]], desc_ref[f, :cnt[f]])


# ---- NN detector post-processing (fd_nn.cu) ----------------------------------------------------------------------------------
def test_nn_postprocessing_vs_checker(ctx, checker, torch_cuda):
    """Heat map -> candidates -> greedy selection -> descriptor sampling against nn_feature_point_detector.cpp compiled in
    place: keypoints identical (equal responses included: the later raster position first), descriptors bit-exact."""
    from feature_detector_b200.synth import synth_descriptor_volume, synth_heatmap
    torch = torch_cuda
    rng = np.random.default_rng(21)
    for trial, (w, h, n_frames) in enumerate([(752, 480, 3), (333, 217, 4), (160, 120, 6), (64, 40, 2)]):
        for q in (0.0, 1.0 / 32):
            maps = np.stack([synth_heatmap(w, h, 40 + trial * 10 + f, q) for f in range(n_frames)])
            d_maps = torch.from_numpy(maps).cuda()
            pres = [np.stack([rng.integers(0, w, 5 + f), rng.integers(0, h, 5 + f)], 1).astype(np.float32) for f in range(n_frames)]
            for thr, b, dist, n, with_pre in ((0.1, 3, 15, 240, False), (0.05, 3, 9, 120, True), (0.3, 0, 2, 1000, False), (0.02, 9, 40, 5, True),
                                              (0.1, 3, 15, 0, False)):
                ctx.set_existing_features(pres if with_pre else [])
                ctx.nn_select(d_maps.data_ptr(), h, w, n_frames, fd.NnParams(thr, b, dist, n), 0)
                kp, cnt = ctx.keypoints(max(n, 1))
                feats = []
                for f in range(n_frames):
                    o = checker.nn_select(maps[f], thr, b, dist, n, pres[f] if with_pre else None)
                    new = o["features"][len(pres[f]) if with_pre else 0:]
                    got = np.stack([kp["x"][f, :cnt[f]], kp["y"][f, :cnt[f]]], 1).astype(np.float32)
                    assert np.array_equal(got, new), (trial, q, thr, f)
                    assert np.array_equal(kp["response"][f, :cnt[f]], maps[f][got[:, 1].astype(int), got[:, 0].astype(int)])
                    feats.append(got)
                for ch in (256, 128):
                    vol = np.stack([synth_descriptor_volume(ch, h // 8, w // 8, trial + f) for f in range(n_frames)])
                    d_vol = torch.from_numpy(vol).cuda()
                    ctx.nn_sample_descriptors(d_vol.data_ptr(), ch, h // 8, w // 8)
                    desc = ctx.nn_descriptors(max(n, 1))
                    for f in range(n_frames):
                        exp = checker.nn_descriptors(feats[f], vol[f])
                        assert np.array_equal(desc[f, :cnt[f]].view(np.uint32), exp.view(np.uint32)), (trial, q, thr, f, ch)
                    if with_pre:   # the reference describes the pre-existing features too; fractional and border points included
                        pts = [np.concatenate([pres[f], np.array([[0, 0], [w - 1, h - 1], [3.5, 7.25], [w - 9, 2]], np.float32)]) for f in range(n_frames)]
                        got = ctx.nn_descriptors_at(d_vol.data_ptr(), ch, h // 8, w // 8, pts)
                        for f in range(n_frames):
                            assert np.array_equal(got[f].view(np.uint32), checker.nn_descriptors(pts[f], vol[f]).view(np.uint32)), (trial, f, ch)
    ctx.set_existing_features([])


def test_nn_postprocessing_golden(ctx, torch_cuda):
    import os
    import importlib.util
    from feature_detector_b200.synth import synth_descriptor_volume, synth_heatmap
    here = os.path.dirname(__file__)
    spec = importlib.util.spec_from_file_location("make_golden_nn", os.path.join(here, "golden", "make_golden_nn.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    gold = np.load(os.path.join(here, "golden", "nn_vectors.npz"))
    torch = torch_cuda
    for name, w, h, idx, q, thr, b, d, n, n_pre, ch in mod.CASES:
        hm = synth_heatmap(w, h, idx, q)
        pre = mod.pre_features(w, h, n_pre, idx) if n_pre else None
        ctx.set_existing_features([pre] if n_pre else [])
        d_hm = torch.from_numpy(hm[None]).cuda()
        ctx.nn_select(d_hm.data_ptr(), h, w, 1, fd.NnParams(thr, b, d, n), 0)
        kp, cnt = ctx.keypoints(n)
        got = np.stack([kp["x"][0, :cnt[0]], kp["y"][0, :cnt[0]]], 1).astype(np.float32)
        assert np.array_equal(got, gold[name + ".features"][n_pre:]), name
        vol = synth_descriptor_volume(ch, h // 8, w // 8, idx)
        d_vol = torch.from_numpy(vol[None]).cuda()
        ctx.nn_sample_descriptors(d_vol.data_ptr(), ch, h // 8, w // 8)
        desc = ctx.nn_descriptors(n)[0, :cnt[0]]
        # the golden checksum covers the pre-existing features' descriptors too: sample those through the same kernel path
        if n_pre == 0:
            assert np.array_equal(desc[:4].view(np.uint32), gold[name + ".desc_first"].view(np.uint32)), name
            assert np.array_equal(desc.astype(np.float64).sum(0), gold[name + ".desc_sum"]), name
    ctx.set_existing_features([])


# ---- races show up as run-to-run differences: the same batch, repeatedly and under different work distributions ---------------
@pytest.mark.gpu
def test_results_do_not_depend_on_scheduling(monkeypatch):
    """Work items are handed out by global counters and candidates are appended with atomics, so the ORDER inside the candidate
    slots varies from run to run; everything a caller can see (sorted candidates, keypoints, descriptors, LSD maps and seed order)
    must not.  Three repeats per context, and contexts whose bands are cut differently (FD_B200_ITEMS_PER_WARP 1 / 8 / 37)."""
    from feature_detector_b200.synth import synth
    frames = np.stack([synth(333, 217, i) for i in range(12)])
    big = np.stack([synth(752, 480, 100 + i) for i in range(4)])

    def snapshot(c):
        out = []
        for batch in (frames, big):
            c.upload(batch)
            for prm in (fd.DetectParams(fd.FAST, 10.0, 20, 200, fast_n=9), fd.DetectParams(fd.FAST, 10.0, 20, 200, fast_n=12),
                        fd.DetectParams(fd.FAST, 0.1, 15, 100, fast_n=12), fd.DetectParams(fd.HARRIS, 30.0, 20, 200),
                        fd.DetectParams(fd.SHI_TOMAS, 40.0, 20, 300)):
                c.detect(prm)
                c.describe_selected(fd.BriefParams(256, 8))
                n = int(prm.needed_feature_num)
                kp, cnt = c.keypoints(n)
                desc = c.descriptors(n)
                out.append(cnt.copy())
                out.append(c.candidate_counts())
                for f in range(len(batch)):
                    out.append(kp[f, :cnt[f]].copy())
                    out.append(desc[f, :cnt[f]].copy())
                out.append(_cand_table(c, 0))
            c.lsd_field(fd.LsdParams(20.0, 1))
            for f in (0, len(batch) - 1):
                g = c.lsd_download(f)
                out += [g["norm"], g["angle"], g["sorted_idx"]]
        return out

    def same(a, b):
        return len(a) == len(b) and all(x.dtype == y.dtype and x.shape == y.shape and x.tobytes() == y.tobytes() for x, y in zip(a, b))

    base = None
    for items in ("8", "1", "37"):
        monkeypatch.setenv("FD_B200_ITEMS_PER_WARP", items)
        with fd.Context(0) as c:
            for rep in range(3):
                snap = snapshot(c)
                if base is None:
                    base = snap
                assert same(base, snap), (items, rep)


# ---- Hamming matching of the packed descriptors (fd_match.cu; SURVEY.md 8f-4 -- the reference has no matcher, numpy is the checker) ----
def _numpy_matches(desc, counts, capacity):
    """For every descriptor of frame f the nearest one of frame f + 1 by Hamming distance (lowest index on ties), that distance and the
    second-smallest distance -- XOR + popcount over the 32 bytes."""
    lut = np.array([bin(i).count("1") for i in range(256)], np.int32)
    out = np.zeros((len(desc) - 1, capacity), fd.MATCH_DTYPE)
    out["train_index"], out["distance"], out["second_distance"] = -1, -1, -1
    for f in range(len(desc) - 1):
        a, b = desc[f, :counts[f]], desc[f + 1, :counts[f + 1]]
        if len(a) == 0 or len(b) == 0:
            continue
        dist = lut[a[:, None, :] ^ b[None, :, :]].sum(-1)
        idx = dist.argmin(1)
        out["train_index"][f, :len(a)] = idx
        out["distance"][f, :len(a)] = dist[np.arange(len(a)), idx]
        if len(b) > 1:
            out["second_distance"][f, :len(a)] = np.sort(dist, 1)[:, 1]
    return out


def test_hamming_matches_vs_numpy(ctx, torch_cuda):
    from feature_detector_b200.synth import synth
    torch = torch_cuda
    frames = np.stack([synth(752, 480, 300 + i // 2) for i in range(7)])        # pairs of equal frames among them: distance 0 matches
    frames[3] = 0                                                                # a frame without keypoints: no matches into or out of it
    ctx.upload(frames)
    ctx.set_existing_features([])
    ctx.detect(fd.DetectParams(fd.FAST, 10.0, 20, 150, fast_n=9))
    ctx.describe_selected(fd.BriefParams(256, 8))
    kp, cnt = ctx.keypoints(150)
    desc = ctx.descriptors(150)
    ctx.match_selected()
    got = ctx.matches(150)
    want = _numpy_matches(desc, cnt, 150)
    for name in ("train_index", "distance", "second_distance"):
        assert np.array_equal(got[name], want[name]), name
    same = got["distance"][0, :cnt[0]]
    nonzero = desc[0, :cnt[0]].any(-1)                                          # (keypoints inside the border get all-zero descriptors, which match each other)
    assert cnt[0] > 50 and np.all(same == 0) and np.array_equal(got["train_index"][0, :cnt[0]][nonzero], np.arange(cnt[0])[nonzero])   # frame 0 == frame 1
    assert np.all(got["train_index"][2] == -1) and np.all(got["train_index"][3] == -1)
    # caller-supplied sets: random descriptors, different capacities
    rng = np.random.default_rng(3)
    a = torch.from_numpy(rng.integers(0, 256, (5, 40, 32), dtype=np.uint8)).cuda()
    b = torch.from_numpy(rng.integers(0, 256, (5, 333, 32), dtype=np.uint8)).cuda()
    ca = torch.tensor([40, 0, 7, 40, 1], dtype=torch.int32).cuda()
    cb = torch.tensor([333, 10, 0, 1, 200], dtype=torch.int32).cuda()
    out = torch.zeros((5, 40, 4), dtype=torch.int32).cuda()
    torch.cuda.synchronize()
    ctx.match_descriptors(a.data_ptr(), ca.data_ptr(), 40, b.data_ptr(), cb.data_ptr(), 333, 5, out.data_ptr())
    ctx.sync()
    o = out.cpu().numpy()
    lut = np.array([bin(i).count("1") for i in range(256)], np.int32)
    an, bn = a.cpu().numpy(), b.cpu().numpy()
    for p_ in range(5):
        na, nb = int(ca[p_]), int(cb[p_])
        for q in range(40):
            if q >= na or nb == 0:
                assert o[p_, q, 0] == -1 and o[p_, q, 1] == -1
                continue
            dist = lut[an[p_, q][None, :] ^ bn[p_, :nb]].sum(-1)
            assert o[p_, q, 0] == dist.argmin() and o[p_, q, 1] == dist.min()
            assert o[p_, q, 2] == (np.sort(dist)[1] if nb > 1 else -1)


def test_one_call_host_to_host_equals_the_separate_calls(ctx, checker):
    """fd_detect_describe_host (upload, detect, describe, downloads with one synchronisation) == the separate calls, for one frame (the
    reference's call pattern), a few frames, pre-existing features, no descriptors, a host capacity below the device's, and an overflow."""
    from feature_detector_b200.synth import synth
    frames = np.stack([synth(752, 480, 500 + i) for i in range(3)])
    det, brief = fd.DetectParams(fd.FAST, 10.0, 20, 200, fast_n=9), fd.BriefParams(256, 8)
    for batch in (frames[:1], frames):
        for existing in (None, [np.array([[100.0, 100.0], [400.0, 300.0]], np.float32)] * len(batch)):
            ctx.set_existing_features(existing or [])
            ctx.upload(batch)
            ctx.detect(det)
            ctx.describe_selected(brief)
            kp_ref, cnt_ref = ctx.keypoints(200)
            desc_ref = ctx.descriptors(200)
            kp, cnt, desc = ctx.detect_describe_host(batch, det, brief, 200)
            assert np.array_equal(cnt, cnt_ref)
            for f in range(len(batch)):
                assert np.array_equal(kp[f, :cnt[f]], kp_ref[f, :cnt[f]]) and np.array_equal(desc[f, :cnt[f]], desc_ref[f, :cnt[f]])
            with pytest.raises(fd.FdError):                       # a host buffer smaller than a frame's keypoint count is an error, as in fd_download_keypoints
                ctx.detect_describe_host(batch, det, brief, int(cnt_ref.max()) - 1)
            kp, cnt, desc = ctx.detect_describe_host(batch, det, None, 200)
            assert desc is None and np.array_equal(cnt, cnt_ref) and np.array_equal(kp[0, :cnt[0]], kp_ref[0, :cnt[0]])
    ctx.set_existing_features([])
    # results of more than 1 MB leave by strided copies instead of the packed block
    many, det_many = np.concatenate([frames] * 8), fd.DetectParams(fd.FAST, 10.0, 5, 1000, fast_n=9)
    ctx.upload(many)
    ctx.detect(det_many)
    ctx.describe_selected(brief)
    (kp_ref, cnt_ref), desc_ref = ctx.keypoints(1000), ctx.descriptors(1000)
    kp, cnt, desc = ctx.detect_describe_host(many, det_many, brief, 1000)
    assert len(many) * 1000 * 48 > (1 << 20) and np.array_equal(cnt, cnt_ref) and cnt.max() > 200
    for f in range(len(many)):
        assert np.array_equal(kp[f, :cnt[f]], kp_ref[f, :cnt[f]]) and np.array_equal(desc[f, :cnt[f]], desc_ref[f, :cnt[f]])
    o = checker.detect(FAST, frames[0], 10.0, 20, 200, fast_n=9)
    kp, cnt, _ = ctx.detect_describe_host(frames[0], det, brief, 200)
    assert np.array_equal(np.stack([kp["x"][0, :cnt[0]], kp["y"][0, :cnt[0]]], 1), o["features"])
    with pytest.raises(fd.FdError):
        ctx.detect_describe_host(frames[0], fd.DetectParams(fd.FAST, 0.1, 15, 200), brief, 200, cand_capacity=1000)


def test_prepared_first_rank_range_equals_in_kernel_one(checker):
    """With few frames and many candidates the rank histogram and the first rank range are prepared by select_hist_kernel /
    select_admit_kernel (many CTAs per frame); FD_B200_SELECT_PREPARE=0 makes the selection kernel stream over its frame itself.
    Same keypoints either way -- both selection forms, one and several rank ranges, masks, and against the checker."""
    import os
    from feature_detector_b200.synth import synth
    os.environ["FD_B200_SELECT_PREPARE"] = "0"
    try:
        plain = fd.Context(0)
    finally:
        del os.environ["FD_B200_SELECT_PREPARE"]
    prepared = fd.Context(0)
    try:
        rng = np.random.default_rng(23)
        batches = [np.stack([synth(752, 480, 40 + i) for i in range(3)]), synth(1920, 1080, 3)[None], rng.integers(0, 256, (2, 300, 400), dtype=np.uint8)]
        cases = [(fd.HARRIS, 30.0, 20, 200, 12), (fd.HARRIS, 0.1, 5, 3000, 12), (fd.FAST, 0.1, 15, 200, 12), (fd.FAST, 0.1, 15, 5000, 9),
                 (fd.SHI_TOMAS, 0.1, 20, 50, 12), (fd.FAST, 10.0, 20, 200, 9)]
        for frames in batches:
            h, w = frames.shape[1:]
            existing = [np.stack([rng.integers(0, w, 9), rng.integers(0, h, 9)], 1).astype(np.float32) for _ in range(len(frames))]
            for kind, thr, d, n, fast_n in cases:
                for with_existing in (False, True):
                    out = []
                    for c in (plain, prepared):
                        c.upload(frames)
                        c.set_existing_features(existing if with_existing else [])
                        c.detect(fd.DetectParams(kind, thr, d, n, fast_n=fast_n), 0)
                        out.append(c.keypoints(max(n, 1)))
                    assert np.array_equal(out[0][1], out[1][1]), (kind, thr, d, n, with_existing)
                    for f in range(len(frames)):
                        k = out[0][1][f]
                        assert np.array_equal(out[0][0][f, :k], out[1][0][f, :k]), (kind, thr, d, n, with_existing, f)
        prepared.set_existing_features([])
        prepared.upload(batches[1])
        prepared.detect(fd.DetectParams(fd.FAST, 0.1, 15, 300, fast_n=12), 0)
        kp, cnt = prepared.keypoints(300)
        o = checker.detect(FAST, batches[1][0], 0.1, 15, 300, fast_n=12)
        assert np.array_equal(np.stack([kp["x"][0, :cnt[0]], kp["y"][0, :cnt[0]]], 1).astype(np.float32), o["features"])
    finally:
        plain.close()
        prepared.close()


def test_describe_selected_refuses_keypoints_of_other_frames(ctx, torch_cuda):
    """Keypoints selected from caller-supplied candidate keys (row tiles, gathered keys) or from a heat map do not belong to the bound
    frames -- another frame count, another coordinate frame -- so fd_describe_selected answers FD_ERR_NOT_READY instead of reading the
    keypoint slots of a different selection out of bounds."""
    from feature_detector_b200.synth import synth
    torch = torch_cuda
    frames = np.stack([synth(320, 200, i) for i in range(3)])
    ctx.upload(frames)
    ctx.set_existing_features([])
    prm = fd.DetectParams(fd.HARRIS, 30.0, 20, 50)
    ctx.detect(prm)
    ctx.describe_selected(fd.BriefParams(256, 8))                      # fine: the selection covered the bound frames
    keys_ptr, counts_ptr, cap = ctx.device_candidates()
    ctx.select_candidates(prm, keys_ptr, counts_ptr, cap, 200, 320, 1)  # one frame's worth of external keys
    with pytest.raises(fd.FdError) as e:
        ctx.describe_selected(fd.BriefParams(256, 8))
    assert e.value.status == 6                                          # FD_ERR_NOT_READY
    ctx.detect(prm)
    ctx.describe_selected(fd.BriefParams(256, 8))                      # and usable again after a detect over the bound frames


@pytest.mark.parametrize("shape", [(7, 65535), (65535, 7), (9, 40000), (40000, 9), (23, 16389)])
def test_extreme_shapes_vs_checker(ctx, checker, shape):
    """Frames at the coordinate limit of the candidate keys (65535 rows or columns: 16 bits each) and with extreme aspect ratios --
    row pitches that are and are not multiples of 16 (the TMA paths need the former), thousands of column strips, bands of a few
    rows -- against the checker: candidates, keypoints, the LSD field and BRIEF descriptors near the far corner."""
    h, w = shape
    rng = np.random.default_rng(h * 131 + w)
    yy, xx = np.mgrid[0:h, 0:w]
    im = (96 + 64 * (((yy // 3) + (xx // 5)) % 2) + rng.integers(0, 24, (h, w))).astype(np.uint8)
    for name, thr, d, n, fast_n in (("fast", 10.0, 20, 200, 9), ("harris", 30.0, 20, 200, 12), ("shi", 40.0, 3, 500, 12)):
        o = checker.detect(KIND[name], im, thr, d, n, fast_n=fast_n)
        feats, cand, kp = _gpu_detect(ctx, im, KIND[name], thr, d, n, fast_n)
        _assert_same_candidates(cand, o)
        _assert_same_features(feats, cand, o, im.shape, d, n)
        assert o["n_cand"] > 0, (name, shape)
    # descriptors of the keypoints closest to the last row / column
    ctx.describe_selected(fd.BriefParams(256, 8))
    bits = fd.unpack_bits(ctx.descriptors(500)[0, :len(feats)], 256)
    far = np.argsort(-(feats[:, 0] + feats[:, 1]))[:20]
    ok, exp = checker.brief(im, feats[far], 256, 8)
    assert ok and np.array_equal(bits[far], exp)
    ctx.upload(im)
    ctx.lsd_field(fd.LsdParams(20.0, 1))
    g = ctx.lsd_download(0)
    o = checker.lsd_map(im)
    assert np.array_equal(g["norm"][:h - 1, :w - 1].view(np.uint32), o["norm"].view(np.uint32))
    valid = g["norm"][:h - 1, :w - 1] > 20.0
    assert np.array_equal(valid, o["valid"].astype(bool)) and g["n_valid"] == int(o["valid"].sum())
    assert np.max(np.abs(g["angle"][:h - 1, :w - 1][valid] - o["angle"][valid]), initial=0.0) <= 1e-5
    exp_rc = o["sorted_rc"]
    assert np.array_equal(g["norm"].reshape(-1)[g["sorted_idx"]], o["norm"][exp_rc[:, 0], exp_rc[:, 1]])
