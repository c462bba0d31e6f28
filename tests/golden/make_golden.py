#!/usr/bin/env python
"""Generate the committed golden vectors from the reference itself (oracle/_ref/libfd_ref.so).

Runs only where /root/reference is mounted (this container).  Outputs, all small:
  tests/golden/image_752x480.u8.gz   raw decode (cv2.IMREAD_GRAYSCALE) of reference examples/image.png
  tests/golden/kat.json              counts / sums / hashes per case (regenerated)
  tests/golden/vectors.npz           keypoint lists, packed BRIEF descriptors, LSD seed norms for image.png and
                                     two small synthetic frames

Hash = 64-bit FNV-1a with the survey's (truncated) offset basis 1469598103934665603, so that the numbers
can be compared with SURVEY.md directly.
"""
import gzip
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from feature_detector_b200.synth import synth  # noqa: E402
from oracle.bindings import FAST, HARRIS, SHI_TOMAS, Port, Ref  # noqa: E402
from oracle.tiecheck import greedy_replay, same_up_to_ties  # noqa: E402

SURVEY_BASIS = 1469598103934665603


def fnv(arr, basis=SURVEY_BASIS):
    h = basis
    for b in np.ascontiguousarray(arr).tobytes():
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return "%016x" % h


def cand_record(o):
    xy, r = o["cand_xy"], o["cand_resp"]
    order = np.lexsort((xy[:, 0], xy[:, 1]))
    flat = np.zeros((len(r), 3), np.uint32)
    flat[:, 0] = r[order].view(np.uint32)
    flat[:, 1] = xy[order, 0]
    flat[:, 2] = xy[order, 1]
    return flat


def main():
    import cv2
    ref = Ref()
    port = Port()
    img = cv2.imread("/root/reference/examples/image.png", cv2.IMREAD_GRAYSCALE)
    assert img.shape == (480, 752)
    with gzip.GzipFile(os.path.join(HERE, "image_752x480.u8.gz"), "wb", mtime=0) as f:
        f.write(img.tobytes())
    frames = {"image": img, "synth752": synth(752, 480, 0), "synth_odd": synth(333, 217, 5)}
    kat = {"image_sha256": hashlib.sha256(img.tobytes()).hexdigest(), "hash": "fnv1a64 basis %d" % SURVEY_BASIS, "cases": []}
    vec = {}
    cases = [("fast", FAST, 10.0, 20, 200, 0), ("fast", FAST, 0.1, 15, 200, 0), ("fast9", FAST, 10.0, 20, 200, 9),
             ("harris", HARRIS, 30.0, 20, 200, 0), ("harris", HARRIS, 0.1, 15, 200, 0),
             ("shi", SHI_TOMAS, 40.0, 20, 200, 0), ("shi", SHI_TOMAS, 0.1, 15, 1000, 0)]
    for fname, im in frames.items():
        for name, kind, thr, d, n, fn in cases:
            o = ref.detect(kind, im, thr, d, n, fast_n=fn, want_response=(kind != FAST))
            rec = cand_record(o)
            ties = int(len(rec) - len(np.unique(rec[:, 0])))
            case = {"frame": fname, "detector": name, "thr": thr, "dist": d, "needed": n, "fast_n": fn or 12, "n_cand": int(o["n_cand"]),
                    "cand_hash": fnv(rec), "cand_sum": float(o["cand_resp"].astype(np.float64).sum()), "n_feat": int(len(o["features"])),
                    "feat_hash": fnv(o["features"].astype("<i4")), "tied_candidates": ties}
            if kind != FAST:
                case["resp_hash"] = fnv(o["response"])
                case["resp_nonzero"] = int(np.count_nonzero(o["response"]))
            # The reference's std::sort is unstable; this framework breaks ties in raster order (the port does too).
            # When that changes the selection, verify the two are identical except for ties and record both.
            q = port.detect(kind, im, thr, d, n, fast_n=fn)
            case["tie_sensitive"] = not np.array_equal(q["features"], o["features"])
            if case["tie_sensitive"]:
                assert same_up_to_ties(o["cand_resp"], o["cand_xy"], q["cand_resp"], q["cand_xy"])
                for res in (o, q):
                    replay, _ = greedy_replay(res["cand_xy"], im.shape[0], im.shape[1], d, n)
                    assert np.array_equal(replay, res["features"])
            case["feat_hash_raster_ties"] = fnv(q["features"].astype("<i4"))
            case["n_feat_raster_ties"] = int(len(q["features"]))
            kat["cases"].append(case)
            vec[f"{fname}.{name}.{thr:g}.{d}.{n}.features"] = q["features"].astype(np.int16)
        for fn in (12, 9):
            s = ref.fast_score_map(im, fn, 15)
            kat["cases"].append({"frame": fname, "detector": "fast_score", "fast_n": fn, "hist": np.bincount(s[3:-3, 3:-3].ravel(), minlength=17).tolist(),
                                 "score_hash": fnv(s)})
        # BRIEF on Harris keypoints (thr 20, d 20, N 200) + a few fractional keypoints
        kp = ref.detect(HARRIS, im, 20.0, 20, 200, want_candidates=False)["features"]
        rng = np.random.default_rng(7)
        h, w = im.shape
        frac = np.stack([rng.uniform(0, w, 40), rng.uniform(0, h - 20, 40)], 1).astype(np.float32)
        for tag, pts, length in (("harris200", kp, 256), ("harris10", kp[:10], 128), ("frac40", frac, 256)):
            ok, bits = ref.brief(im, pts, length, 8)
            packed = np.packbits(bits, axis=1, bitorder="little")
            kat["cases"].append({"frame": fname, "detector": "brief", "set": tag, "length": length, "n": int(len(pts)), "ones": int(bits.sum()),
                                 "all_zero": int((bits.sum(1) == 0).sum()), "hash": fnv(packed)})
            vec[f"{fname}.brief.{tag}.kp"] = pts.astype(np.float32)
            vec[f"{fname}.brief.{tag}.desc"] = packed
        lsd = ref.lsd_map(im)
        srt = lsd["sorted_rc"]
        kat["cases"].append({"frame": fname, "detector": "lsd", "n_valid": int(lsd["valid"].sum()), "norm_hash": fnv(lsd["norm"]),
                             "angle_hash": fnv(lsd["angle"]), "norm_sum": float(lsd["norm"].astype(np.float64).sum()),
                             "angle_sum": float(lsd["angle"].astype(np.float64).sum()),
                             "sorted_norm_hash": fnv(lsd["norm"][srt[:, 0], srt[:, 1]])})
    # pre-seeded Harris case of the demo (test_feature_point_detector.cpp:48-57)
    pre = np.array([[15 * i, 15 * j] for i in range(1, 10) for j in range(1, 10)], np.float32)
    o = ref.detect(HARRIS, img, 30.0, 20, 200, pre=pre)
    kat["cases"].append({"frame": "image", "detector": "harris_preseeded81", "thr": 30.0, "dist": 20, "needed": 200, "n_cand": int(o["n_cand"]),
                         "n_feat": int(len(o["features"])), "feat_hash": fnv(o["features"].astype("<i4"))})
    vec["image.harris_preseeded81.features"] = o["features"].astype(np.int16)
    # sparsify
    rng = np.random.default_rng(3)
    f = np.stack([rng.uniform(-20, 800, 400), rng.uniform(-20, 520, 400)], 1).astype(np.float32)
    st = rng.integers(0, 3, 400).astype(np.uint8)
    vec["sparsify.features"] = f
    vec["sparsify.status_in"] = st
    vec["sparsify.status_out"] = ref.sparsify(f, 480, 752, 1, 2, st)
    with open(os.path.join(HERE, "kat.json"), "w") as fjs:
        json.dump(kat, fjs, indent=1)
    np.savez_compressed(os.path.join(HERE, "vectors.npz"), **vec)
    print("cases:", len(kat["cases"]), "vectors:", len(vec))


if __name__ == "__main__":
    main()
