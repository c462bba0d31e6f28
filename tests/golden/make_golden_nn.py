#!/usr/bin/env python
"""Generate the committed golden vectors of the NN post-processing from the reference itself (oracle/_ref/libfd_ref.so:
nn_feature_point_detector.cpp compiled in place).  Runs only where /root/reference is mounted.  Output:
tests/golden/nn_vectors.npz -- per case the selected features and a checksum of the sampled descriptors; the inputs are
regenerated from feature_detector_b200.synth (synth_heatmap / synth_descriptor_volume)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from feature_detector_b200.synth import synth_descriptor_volume, synth_heatmap  # noqa: E402
from oracle.bindings import Ref  # noqa: E402

# (name, width, height, idx, quantum, min_response, invalid_boundary, min_distance, max_features, n_pre, channels)
CASES = [("defaults", 752, 480, 0, 0.0, 0.1, 3, 15, 240, 0, 256),
         ("ties", 752, 480, 1, 1.0 / 32, 0.1, 3, 15, 240, 0, 256),
         ("preseeded", 320, 200, 2, 0.0, 0.05, 3, 9, 120, 20, 128),
         ("odd_no_boundary", 333, 217, 3, 1.0 / 16, 0.2, 0, 4, 1000, 7, 256),
         ("wide_boundary", 160, 120, 4, 0.0, 0.01, 11, 30, 50, 0, 128)]


def pre_features(width, height, n, idx):
    rng = np.random.default_rng([width, height, n, idx])
    return np.stack([rng.integers(0, width, n), rng.integers(0, height, n)], 1).astype(np.float32)


def main():
    ref = Ref()
    out = {}
    for name, w, h, idx, q, thr, b, d, n, n_pre, ch in CASES:
        hm = synth_heatmap(w, h, idx, q)
        pre = pre_features(w, h, n_pre, idx) if n_pre else None
        sel = ref.nn_select(hm, thr, b, d, n, pre)
        vol = synth_descriptor_volume(ch, h // 8, w // 8, idx)
        desc = ref.nn_descriptors(sel["features"], vol)
        out[name + ".features"] = sel["features"]
        out[name + ".n_cand"] = np.int64(sel["n_cand"])
        out[name + ".desc_sum"] = desc.astype(np.float64).sum(0).astype(np.float64)
        out[name + ".desc_first"] = desc[:4].copy()
        print(name, sel["n_cand"], len(sel["features"]))
    np.savez_compressed(os.path.join(HERE, "nn_vectors.npz"), **out)


if __name__ == "__main__":
    main()
