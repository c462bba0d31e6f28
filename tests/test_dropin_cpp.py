"""The C++ drop-in classes (feature_detector_b200/cpp) replaying the reference's demo sequences.

fd_dropin_check runs test/test_feature_point_detector.cpp, test_feature_descriptor.cpp and the dense stage of
test_feature_line_detector.cpp through the reference-named classes; its output is compared with the golden answers
(regenerated from the reference build by tests/golden/make_golden.py).
"""
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPP = os.path.join(ROOT, "feature_detector_b200", "cpp")
EXE = os.path.join(CPP, "fd_dropin_check")


NN_CHANNELS, NN_PRE = 256, 5


def _nn_inputs():
    from feature_detector_b200.synth import synth_descriptor_volume, synth_heatmap
    return synth_heatmap(752, 480, 7), synth_descriptor_volume(NN_CHANNELS, 60, 94, 7)


@pytest.fixture(scope="module")
def dropin_output(built, image_png, tmp_path_factory):
    r = subprocess.run(["make", "-C", CPP], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-3000:]
    return _run_check(EXE, image_png, tmp_path_factory.mktemp("dropin"))


def _run_check(exe, image_png, tmp):
    raw = tmp / "image_752x480.u8"
    raw.write_bytes(image_png.tobytes())
    heat, vol = _nn_inputs()
    (tmp / "heat.f32").write_bytes(heat.tobytes())
    (tmp / "vol.f32").write_bytes(vol.tobytes())
    r = subprocess.run([exe, str(raw), "480", "752", str(tmp / "heat.f32"), str(tmp / "vol.f32"), str(NN_CHANNELS), str(NN_PRE)],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    return json.loads(r.stdout)


@pytest.fixture(scope="module")
def dropin_output_host_glue(built, image_png, tmp_path_factory):
    """fd_dropin_check and the drop-in classes linked against tests/hoststage/fake_fd_abi.cpp (the C ABI with the CPU oracle behind it,
    test infrastructure) instead of libfd_b200.so: the host glue of feature_detector_b200/cpp runs without a GPU."""
    build = os.path.join(ROOT, "tests", "_build")
    os.makedirs(build, exist_ok=True)
    exe = os.path.join(build, "fd_dropin_check_host_glue")
    srcs = [os.path.join(CPP, f) for f in ("fd_dropin_check.cpp", "feature_point_detector.cpp", "descriptor_brief.cpp", "feature_line_field.cpp",
                                           "nn_feature_point_postprocess.cpp")] + [os.path.join(ROOT, "tests", "hoststage", "line_segments_host.cpp")]
    cmd = ["g++", "-std=c++17", "-O2", "-g", "-Wall", "-I" + CPP, "-I" + os.path.join(ROOT, "tests", "hoststage"), "-I" + os.path.join(ROOT, "include"),
           "-I" + os.path.join(ROOT, "compat", "slam_utility"),
           "-o", exe] + srcs + [os.path.join(ROOT, "tests", "hoststage", "fake_fd_abi.cpp"), "-L" + os.path.join(ROOT, "oracle"), "-lfd_oracle",
                                "-Wl,-rpath," + os.path.join(ROOT, "oracle")]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-4000:]
    return _run_check(exe, image_png, tmp_path_factory.mktemp("dropin_host_glue"))


def _feature_hash(features):
    h = 1469598103934665603
    for b in np.ascontiguousarray(np.asarray(features).astype("<i4")).tobytes():
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return "%016x" % h


def _mask_hash(mask):
    h = 1469598103934665603
    for b in np.ascontiguousarray(mask, "<i4").tobytes():
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return "%016x" % h


def test_dropin_headers_keep_the_reference_surface():
    """Names a caller of the reference compiles against."""
    hdr = open(os.path.join(CPP, "feature_point_detector.h")).read()
    for name in ("namespace feature_detector", "class FeaturePointDetector", "class FeaturePointHarrisDetector", "class FeaturePointShiTomasDetector",
                 "class FeaturePointFastDetector", "bool DetectGoodFeatures(const GrayImage &image, const uint32_t needed_feature_num, std::vector<Vec2> &features)",
                 "void SparsifyFeatures(", "kMinFeatureDistance = 15", "kGridFilterRowDivideNumber = 12", "kMinValidResponse = 0.1f", "Options &options()",
                 "candidates()", "mask()", '"Harris"', '"Shi-Tomas"', '"Fast"', "kN = 12", "kMinPixelDiffValue = 15"):
        assert name in hdr, name
    for inc in ("feature_point_harris_detector.h", "feature_point_shi_tomas_detector.h", "feature_point_fast_detector.h", "descriptor.h", "descriptor_brief.h"):
        assert os.path.exists(os.path.join(CPP, inc)), inc
    brief = open(os.path.join(CPP, "descriptor_brief.h")).read()
    for name in ("using BriefType = std::vector<bool>", "class BriefDescriptor", "kLength = 256", "kHalfPatchSize = 8"):
        assert name in brief, name


def test_dropin_has_no_cpu_path(dropin_output):
    """Without a CUDA device every GPU-backed call returns false; with one, this test is vacuous."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu-marked test")
    out = dropin_output
    assert out["fast_demo"]["ok"] is False and out["harris_demo"]["ok"] is False and out["brief_harris10"]["ok"] is False
    assert out["lsd_field"]["ok"] is False and out["fast_demo"]["n_feat"] == 0 and out["lsd_detect"]["ok"] is False
    assert out["null_image_returns"] is False
    assert out["nn"]["ok"] is False and out["nn"]["n_feat"] == NN_PRE


@pytest.mark.gpu
def test_dropin_nn_postprocessing(dropin_output, checker):
    """NNFeaturePointPostProcessor (heat map and descriptor volume in host memory, five pre-existing features) against the
    reference's nn_feature_point_detector.cpp: same features in the same order, same descriptors bit for bit."""
    _check_nn(dropin_output, checker)


def test_dropin_host_glue_nn_postprocessing(dropin_output_host_glue, checker):
    """The same expectations for the host glue alone (CPU; oracle behind the C ABI)."""
    _check_nn(dropin_output_host_glue, checker)


def _check_nn(dropin_output, checker):
    heat, vol = _nn_inputs()
    pre = np.array([[20 + 31 * i % (752 - 40), 20 + 17 * i % (480 - 40)] for i in range(NN_PRE)], np.float32)
    o = checker.nn_select(heat, 0.1, 3, 15, 240, pre)
    nn = dropin_output["nn"]
    assert nn["ok"] is True and nn["ok_desc"] is True
    assert nn["n_feat"] == len(o["features"]) == nn["n_desc"]
    assert np.array_equal(np.array(nn["features"], np.float32), o["features"])
    exp = checker.nn_descriptors(o["features"], vol)
    h = 1469598103934665603
    for b in np.ascontiguousarray(exp, "<f4").tobytes():
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    assert nn["desc_hash"] == "%016x" % h


@pytest.mark.gpu
def test_dropin_batch_call_equals_single_frame_calls(dropin_output):
    """DetectGoodFeaturesBatch on three frames with per-frame pre-existing features == three DetectGoodFeatures calls, for Harris and for
    FAST at the reference's default threshold (whose candidate count overflows the batch's bounded first slots: the call runs again)."""
    b = dropin_output["batch_of_three"]
    assert b["ok"] is True and b["equals_single_frame_calls"] is True and b["n_features"] > 600, b


@pytest.mark.gpu
def test_dropin_replays_reference_demos(dropin_output, kat, image_png, checker):
    _check_demos(dropin_output, kat, image_png, checker)


def test_dropin_host_glue_replays_reference_demos(dropin_output_host_glue, kat, image_png, checker):
    """fd_dropin_check through the host glue alone (CPU; oracle behind the C ABI): marshalling, the in/out features contract, the
    lazily rebuilt mask() / candidates(), bit unpacking, PixelParam filling and the LSD host stage against the same golden answers."""
    _check_demos(dropin_output_host_glue, kat, image_png, checker)


def _check_demos(out, kat, image_png, checker):
    gold = {(c["detector"], c["thr"]): c for c in kat["cases"] if c["frame"] == "image" and "thr" in c}
    for label, key, name in (("fast_demo", ("fast", 10.0), "Fast"), ("harris_demo", ("harris", 30.0), "Harris"), ("shi_demo", ("shi", 40.0), "Shi-Tomas"),
                             ("fast9_demo", ("fast9", 10.0), "Fast"), ("harris_default_reused", ("harris", 0.1), "Harris")):
        o, g = out[label], gold[key]
        assert o["ok"] is True and o["name"] == name
        assert o["n_cand"] == g["n_cand"], label
        assert o["n_feat"] == g["n_feat_raster_ties"] and o["feat_hash"] == g["feat_hash_raster_ties"], label
        assert (o["mask_rows"], o["mask_cols"]) == (480, 752)
    # SURVEY.md 8c first features
    assert out["fast_demo"]["first"] == [45, 471] and out["harris_demo"]["first"] == [74, 3] and out["shi_demo"]["first"] == [639, 63]
    pre = out["harris_preseeded81"]
    g = [c for c in kat["cases"] if c["detector"] == "harris_preseeded81"][0]
    assert pre["n_pre"] == 81 and pre["n_feat"] == g["n_feat"] == 200 and pre["n_cand"] == g["n_cand"] and pre["feat_hash"] == g["feat_hash"]
    assert pre["first"] == [520, 201]
    # mask(): FAST demo found 84 < 200 features, so every feature cleared its clipped 41x41 square
    assert 0 < out["fast_demo"]["mask_zeros"] <= 84 * 41 * 41
    # mask() after a call is the reference's mask_, exactly: all ones minus the squares of the pre-existing features and of
    # every accepted feature except the one that reached the count (feature_point_detector.cpp:67-69)
    from oracle.bindings import FAST, HARRIS, SHI_TOMAS
    pre81 = np.array([[15 * i, 15 * j] for i in range(1, 10) for j in range(1, 10)], np.float32)
    for label, kind, thr, d, fast_n, pre_feats in (("fast_demo", FAST, 10.0, 20, 12, None), ("harris_demo", HARRIS, 30.0, 20, 0, None),
                                                   ("shi_demo", SHI_TOMAS, 40.0, 20, 0, None), ("fast9_demo", FAST, 10.0, 20, 9, None),
                                                   ("harris_preseeded81", HARRIS, 30.0, 20, 0, pre81), ("harris_default_reused", HARRIS, 0.1, 15, 0, None)):
        o = checker.detect(kind, image_png, thr, d, 200, fast_n=fast_n, pre=pre_feats, want_mask=True, want_candidates=False)
        if len(o["features"]) == out[label]["n_feat"] and np.array_equal(out[label].get("first", [-1, -1]), o["features"][len(pre_feats) if pre_feats is not None else 0].astype(int)):
            assert out[label]["mask_zeros"] == int((o["mask"] == 0).sum()), label
            if out[label]["feat_hash"] == _feature_hash(o["features"]):   # identical feature lists (no tie resolved differently): identical masks
                assert out[label]["mask_hash"] == _mask_hash(o["mask"]), label
    assert out["null_image_returns"] is False

    b = out["brief_harris10"]
    g = [c for c in kat["cases"] if c["frame"] == "image" and c["detector"] == "brief" and c["set"] == "harris10"][0]
    assert b["ok"] is True and b["n"] == 10 and b["length"] == 128
    assert (b["ones"], b["all_zero"], b["hash"]) == (g["ones"], g["all_zero"], g["hash"])
    assert b["float_sum"] == 2.0 * g["ones"] - 10 * 128          # bits -> +1 / -1 (descriptor.h:51-54)
    assert b["empty_returns"] is False                            # descriptor.h:29

    l = out["lsd_field"]
    g = [c for c in kat["cases"] if c["frame"] == "image" and c["detector"] == "lsd"][0]
    assert l["ok"] is True and (l["rows"], l["cols"]) == (479, 751)
    assert l["n_valid"] == l["n_sorted"] == g["n_valid"] and l["norm_hash"] == g["norm_hash"] and l["sorted_norm_hash"] == g["sorted_norm_hash"]
    assert abs(l["norm_sum"] - g["norm_sum"]) < 1e-3 and abs(l["angle_sum"] - g["angle_sum"]) < 1e-5 * g["n_valid"]
    assert l["descending"] is True and l["positions_ok"] is True

    # the whole line detector (dense stage on the GPU, host stage in tests/hoststage/line_segments_host.cpp, a stand-in for the reference's own):
    # 40 segments on image.png, and exactly the reference's where its build is available
    d = out["lsd_detect"]
    assert d["ok"] is True and d["n_lines"] == d["n_rectangles"] == 40 and d["n_seeds"] == 10087
    from oracle.bindings import Ref, have_ref
    if have_ref():
        ok, ref_lines = Ref().lsd_detect(image_png, 200)
        mine = np.array(d["lines"], np.float32).reshape(-1, 4)
        assert ok and mine.shape == ref_lines.shape
        assert np.array_equal(mine, ref_lines)      # same segments, same order, same bits


@pytest.mark.gpu
def test_route_b3_with_the_references_own_host_stage(image_png, tmp_path):
    """INTEGRATION.md B.3 as a program (integration/route_b3.cpp, built into oracle/_ref/ where the reference is mounted): the
    reference's FeatureLineDetector, compiled in place, with ComputeLineLevelAngleMap bound to LineLevelAngleField::Compute +
    FillPixelParams.  On image.png (no equal-norm seed decides a segment there) it must return the reference's segments bit for bit."""
    exe = os.path.join(ROOT, "oracle", "_ref", "fd_route_b3")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/fd_route_b3 was not built (needs /root/reference at build time)")
    from feature_detector_b200.synth import synth
    frames = np.stack([image_png, synth(752, 480, 3)])
    raw = tmp_path / "frames.u8"
    frames.tofile(raw)
    r = subprocess.run([exe, str(raw), "480", "752", "2", "200"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["frames_identical_given_the_same_tie_order"] >= 1 and out["mean_lines"] > 10, out
    assert out["route_b3_ms_per_frame"]["FillPixelParams"] > 0
