"""CPU suite, part 1: the oracle is pinned to the reference.

* the committed golden vectors (generated from oracle/_ref, i.e. the reference's own code) must be
  reproduced by the plain-C port -- this runs everywhere, including the GPU box where /root/reference
  does not exist;
* where oracle/_ref is present, port and reference are also compared directly on more frames.
"""
import os

import numpy as np
import pytest

from conftest import fnv
from oracle.bindings import FAST, HARRIS, SHI_TOMAS
from oracle.tiecheck import greedy_replay, same_up_to_ties

KINDS = {"fast": FAST, "fast9": FAST, "harris": HARRIS, "shi": SHI_TOMAS}


def _cand_record(o):
    xy, r = o["cand_xy"], o["cand_resp"]
    order = np.lexsort((xy[:, 0], xy[:, 1]))
    flat = np.zeros((len(r), 3), np.uint32)
    flat[:, 0] = r[order].view(np.uint32)
    flat[:, 1] = xy[order, 0]
    flat[:, 2] = xy[order, 1]
    return flat


def test_golden_image_is_the_reference_example(image_png, kat):
    import hashlib
    assert hashlib.sha256(image_png.tobytes()).hexdigest() == kat["image_sha256"]
    # SURVEY.md section 8c quotes this digest for examples/image.png
    assert kat["image_sha256"] == "7aed1139f1833cd6f600bad397066109d0a578a7535f7178dd5cab45f90d2190"


def test_survey_known_answers_are_in_the_golden_file(kat):
    """The survey's table (section 8c) was produced independently; the regenerated goldens must agree."""
    want = {("fast", 10.0): (962, "14b52e7f58f59c8f", 84, "dd67c5129a1e7e40"),
            ("fast", 0.1): (343669, "e479cac1f25d041f", 200, "80b9d39bf6d4ded1"),
            ("harris", 30.0): (12894, "dc38c516de0fade7", 200, "32c324070cbd3b8b"),
            ("harris", 0.1): (14542, "801265c5432c71c5", 200, "97927df081f7973b"),
            ("shi", 40.0): (3742, "e2785ec440d6c9a5", 200, "73062c834d19eb5b"),
            ("shi", 0.1): (13123, "474475b49489e281", 591, "7fba4523e43d10c4")}
    seen = 0
    for c in kat["cases"]:
        key = (c.get("detector"), c.get("thr"))
        if c.get("frame") == "image" and key in want:
            n_cand, ch, n_feat, fh = want[key]
            assert (c["n_cand"], c["cand_hash"], c["n_feat"], c["feat_hash"]) == (n_cand, ch, n_feat, fh)
            seen += 1
    assert seen == len(want)


def test_port_reproduces_golden_detector_cases(port, frames, kat, vectors):
    n = 0
    for c in kat["cases"]:
        if c["detector"] not in KINDS:
            continue
        im = frames[c["frame"]]
        o = port.detect(KINDS[c["detector"]], im, c["thr"], c["dist"], c["needed"], fast_n=c["fast_n"], want_response=c["detector"] in ("harris", "shi"))
        assert o["n_cand"] == c["n_cand"]
        assert fnv(_cand_record(o)) == c["cand_hash"]
        if "resp_hash" in c:
            assert fnv(o["response"]) == c["resp_hash"]
            assert int(np.count_nonzero(o["response"])) == c["resp_nonzero"]
        # features: raster tie rule (port) -- equals the reference's list unless the case is tie sensitive
        assert fnv(o["features"].astype("<i4")) == c["feat_hash_raster_ties"]
        if not c["tie_sensitive"]:
            assert fnv(o["features"].astype("<i4")) == c["feat_hash"]
        key = f"{c['frame']}.{c['detector']}.{c['thr']:g}.{c['dist']}.{c['needed']}.features"
        assert np.array_equal(vectors[key].astype(np.float32), o["features"])
        n += 1
    assert n >= 21


def test_port_reproduces_golden_scores_brief_lsd(port, frames, kat, vectors):
    for c in kat["cases"]:
        im = frames[c["frame"]]
        if c["detector"] == "fast_score":
            s = port.fast_score_map(im, c["fast_n"], 15)
            assert fnv(s) == c["score_hash"]
            assert np.bincount(s[3:-3, 3:-3].ravel(), minlength=17).tolist() == c["hist"]
        elif c["detector"] == "brief":
            kp = vectors[f"{c['frame']}.brief.{c['set']}.kp"]
            ok, bits = port.brief(im, kp, c["length"], 8)
            assert ok and int(bits.sum()) == c["ones"] and int((bits.sum(1) == 0).sum()) == c["all_zero"]
            packed = np.packbits(bits, axis=1, bitorder="little")
            assert fnv(packed) == c["hash"]
            assert np.array_equal(packed, vectors[f"{c['frame']}.brief.{c['set']}.desc"])
        elif c["detector"] == "lsd":
            m = port.lsd_map(im)
            assert int(m["valid"].sum()) == c["n_valid"]
            assert fnv(m["norm"]) == c["norm_hash"]
            assert fnv(m["angle"]) == c["angle_hash"]  # same glibc atan2f as the generator
            s = m["sorted_rc"]
            assert fnv(m["norm"][s[:, 0], s[:, 1]]) == c["sorted_norm_hash"]


def test_port_preseeded_and_sparsify_golden(port, image_png, kat, vectors):
    c = [c for c in kat["cases"] if c["detector"] == "harris_preseeded81"][0]
    pre = np.array([[15 * i, 15 * j] for i in range(1, 10) for j in range(1, 10)], np.float32)  # test_feature_point_detector.cpp:52-56
    o = port.detect(HARRIS, image_png, c["thr"], c["dist"], c["needed"], pre=pre)
    assert o["n_cand"] == c["n_cand"] and len(o["features"]) == c["n_feat"]
    assert fnv(o["features"].astype("<i4")) == c["feat_hash"] == "85bd4f625b5448b3"
    st = port.sparsify(vectors["sparsify.features"], 480, 752, 1, 2, vectors["sparsify.status_in"])
    assert np.array_equal(st, vectors["sparsify.status_out"])


def test_port_equals_reference_on_more_frames(port, ref):
    from feature_detector_b200.synth import synth
    for (w, h, idx) in [(752, 480, 11), (640, 360, 2), (97, 61, 4), (1280, 720, 1)]:
        im = synth(w, h, idx)
        for kind, thr, d, n, fn in [(FAST, 10, 20, 200, 0), (FAST, 10, 20, 200, 9), (FAST, 0.1, 15, 50, 0), (HARRIS, 30, 20, 200, 0),
                                    (SHI_TOMAS, 40, 20, 1000, 0), (HARRIS, 0.1, 3, 5000, 0)]:
            a = ref.detect(kind, im, thr, d, n, fast_n=fn, want_response=True, want_mask=True)
            b = port.detect(kind, im, thr, d, n, fast_n=fn, want_response=True, want_mask=True)
            assert np.array_equal(_cand_record(a), _cand_record(b))
            assert np.array_equal(a["response"].view(np.uint32), b["response"].view(np.uint32))
            if np.array_equal(a["features"], b["features"]):
                assert np.array_equal(a["mask"], b["mask"])
            else:  # identical except for ties
                assert same_up_to_ties(a["cand_resp"], a["cand_xy"], b["cand_resp"], b["cand_xy"])
                for res in (a, b):
                    replay, _ = greedy_replay(res["cand_xy"], h, w, d, n)
                    assert np.array_equal(replay, res["features"])
        for fn in (12, 9):
            assert np.array_equal(ref.fast_score_map(im, fn, 15), port.fast_score_map(im, fn, 15))
        kp = ref.detect(HARRIS, im, 20, 20, 100, want_candidates=False)["features"]
        if len(kp):
            assert np.array_equal(ref.brief(im, kp)[1], port.brief(im, kp)[1])
        a, b = ref.lsd_map(im), port.lsd_map(im)
        for k in ("norm", "angle"):
            assert np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32))
        assert np.array_equal(a["valid"], b["valid"])
        sa, sb = a["sorted_rc"], b["sorted_rc"]
        assert np.array_equal(a["norm"][sa[:, 0], sa[:, 1]], b["norm"][sb[:, 0], sb[:, 1]])


def test_edge_cases_port(port):
    # needed = 0 still yields one feature (feature_point_detector.cpp:67-68); tiny / flat images yield none
    from feature_detector_b200.synth import synth
    im = synth(160, 120, 3)
    assert len(port.detect(HARRIS, im, 30, 20, 0)["features"]) == 1
    flat = np.full((64, 64), 77, np.uint8)
    assert port.detect(HARRIS, flat, 0.1, 15, 100)["n_cand"] == 0
    assert port.detect(FAST, flat, 10, 15, 100)["n_cand"] == 0
    o = port.detect(FAST, flat, 0.001, 15, 100)  # the running offset alone crosses the threshold
    assert o["n_cand"] == 58 * 58 - 99
    for shape in [(4, 4), (6, 9), (7, 7)]:
        tiny = np.zeros(shape, np.uint8)
        for kind in (HARRIS, SHI_TOMAS, FAST):
            assert port.detect(kind, tiny, 10, 5, 10)["n_cand"] == 0


# ---- NN detector post-processing (nn_feature_point_detector.cpp:59-72, 128-155, 163-193) ----------------------------------
def _nn_cases():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_nn", os.path.join(os.path.dirname(__file__), "golden", "make_golden_nn.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.CASES, mod.pre_features


def test_port_reproduces_golden_nn_postprocessing(port):
    """The committed vectors come from the reference's own nn_feature_point_detector.cpp (tests/golden/make_golden_nn.py)."""
    from feature_detector_b200.synth import synth_descriptor_volume, synth_heatmap
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "nn_vectors.npz"))
    cases, pre_features = _nn_cases()
    for name, w, h, idx, q, thr, b, d, n, n_pre, ch in cases:
        hm = synth_heatmap(w, h, idx, q)
        pre = pre_features(w, h, n_pre, idx) if n_pre else None
        sel = port.nn_select(hm, thr, b, d, n, pre)
        assert sel["n_cand"] == int(gold[name + ".n_cand"]), name
        assert np.array_equal(sel["features"], gold[name + ".features"]), name
        desc = port.nn_descriptors(sel["features"], synth_descriptor_volume(ch, h // 8, w // 8, idx))
        assert np.array_equal(desc[:4].view(np.uint32), gold[name + ".desc_first"].view(np.uint32)), name
        assert np.array_equal(desc.astype(np.float64).sum(0), gold[name + ".desc_sum"]), name


def test_port_equals_reference_nn_postprocessing(port, ref):
    from feature_detector_b200.synth import synth_descriptor_volume, synth_heatmap
    rng = np.random.default_rng(3)
    for trial, (w, h) in enumerate([(160, 120), (97, 61), (240, 136), (64, 64)]):
        for q in (0.0, 0.125):
            hm = synth_heatmap(w, h, 10 + trial, q)
            pre = np.stack([rng.integers(0, w, 6), rng.integers(0, h, 6)], 1).astype(np.float32) if trial % 2 else None
            for thr, b, d, n in ((0.1, 3, 15, 240), (0.3, 0, 2, 30), (0.02, 9, 40, 5)):
                a, c = ref.nn_select(hm, thr, b, d, n, pre), port.nn_select(hm, thr, b, d, n, pre)
                assert a["n_cand"] == c["n_cand"] and np.array_equal(a["features"], c["features"]), (trial, q, thr)
            for ch in (256, 128):
                vol = synth_descriptor_volume(ch, h // 8, w // 8, trial)
                pts = np.concatenate([a["features"], np.array([[0, 0], [w - 1, h - 1], [w - 9, 3], [3.5, 7.25]], np.float32)])
                assert np.array_equal(ref.nn_descriptors(pts, vol).view(np.uint32), port.nn_descriptors(pts, vol).view(np.uint32))


def test_brief_vec_overload_port_equals_reference(port, ref, image_png):
    """descriptor.h:43-62: the float overload maps bits to +1 / -1."""
    kp = ref.detect(HARRIS, image_png, 20, 20, 60, want_candidates=False)["features"]
    for length in (256, 100):
        ok_a, a = ref.brief_vec(image_png, kp, length)
        ok_b, b = port.brief_vec(image_png, kp, length)
        assert ok_a and ok_b and np.array_equal(a, b)
        assert np.array_equal(a > 0, ref.brief(image_png, kp, length)[1].astype(bool)) and set(np.unique(a)) <= {-1.0, 1.0}


def test_port_equals_reference_on_random_small_inputs(port, ref):
    """Fuzz of the C restatement against the reference compiled in place: random images from 1 x 1 pixels up (noise, binary, flat, step +
    noise), all three detectors, thresholds from -1 to 1e9, distances 0..100, needed 0..1000, pre-existing features, BRIEF at keypoints
    inside and outside the frame, the LSD maps.  Candidates, responses and masks must agree bit for bit; feature lists too unless two
    candidates tie (the reference's std::sort leaves their order open).  (The reference's line detector is skipped below 3 x 3: on a
    two-column frame it writes out of bounds, feature_line_detector.cpp:64-68.)"""
    from oracle.bindings import FAST, HARRIS, SHI_TOMAS
    rng = np.random.default_rng(20261018)
    for it in range(800):
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
        a = port.detect(kind, img, thr, d, needed, fast_n=fast_n, pre=pre, want_mask=True, want_response=(kind != FAST))
        b = ref.detect(kind, img, thr, d, needed, fast_n=fast_n, pre=pre, want_mask=True, want_response=(kind != FAST))
        assert a["ok"] == b["ok"] and a["n_cand"] == b["n_cand"], case
        if kind != FAST:
            assert np.array_equal(a["response"].view(np.uint32), b["response"].view(np.uint32)), case
        ka, kb = np.lexsort((a["cand_xy"][:, 0], a["cand_xy"][:, 1])), np.lexsort((b["cand_xy"][:, 0], b["cand_xy"][:, 1]))
        assert np.array_equal(a["cand_xy"][ka], b["cand_xy"][kb]) and np.array_equal(a["cand_resp"][ka].view(np.uint32), b["cand_resp"][kb].view(np.uint32)), case
        if np.array_equal(a["features"], b["features"]):
            assert np.array_equal(a["mask"], b["mask"]), case
        else:
            assert len(np.unique(a["cand_resp"])) < len(a["cand_resp"]), case      # only a tie may reorder the walk
        if rows >= 3 and cols >= 3:
            la, lb = port.lsd_map(img), ref.lsd_map(img)
            assert np.array_equal(la["norm"].view(np.uint32), lb["norm"].view(np.uint32)) and np.array_equal(la["valid"], lb["valid"]), case
            assert np.array_equal(la["angle"].view(np.uint32), lb["angle"].view(np.uint32)), case
        if rows > 40 and cols > 40:
            kp = np.stack([rng.uniform(-2, cols + 2, 6), rng.uniform(-2, rows + 2, 6)], 1).astype(np.float32)
            ba, bb = port.brief(img, kp, 256, 8), ref.brief(img, kp, 256, 8)
            assert ba[0] == bb[0] and np.array_equal(ba[1], bb[1]), case
