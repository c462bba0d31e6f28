#!/usr/bin/env python
"""Golden line segments from the reference itself (oracle/_ref/libfd_ref.so: FeatureLineDetector::DetectGoodFeatures compiled in
place).  Runs only where /root/reference is mounted (this container).  Output: tests/golden/lsd_segments.npz -- per case the
(n, 4) float32 segments (start x, start y, end x, end y) in the reference's output order, and `<case>.seeds`: the reference's seed
order as uint16 (row, col) pairs.  The seed order is part of the vector because the reference sorts seeds with an unstable std::sort
(feature_line_detector.cpp:92-94): which of two equal-norm seeds grows first decides the segments, and that is the C++ library's
choice, not the algorithm's."""
import gzip
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from feature_detector_b200.synth import synth  # noqa: E402

# name -> (frame, needed, min gradient norm)
CASES = {
    "image": (None, 200, 20.0),
    "synth752": ((752, 480, 0), 200, 20.0),
    "synth_odd": ((333, 217, 5), 200, 20.0),
    "synth_norm35": ((752, 480, 7), 200, 35.0),
}


def frame_of(name):
    spec = CASES[name][0]
    if spec is None:
        with gzip.open(os.path.join(HERE, "image_752x480.u8.gz"), "rb") as f:
            return np.frombuffer(f.read(), np.uint8).reshape(480, 752)
    return synth(*spec)


if __name__ == "__main__":
    from oracle.bindings import Ref
    ref = Ref()
    out = {}
    for name, (_, needed, min_norm) in CASES.items():
        ok, lines = ref.lsd_detect(frame_of(name), needed, min_norm)
        assert ok
        out[name] = lines.astype(np.float32)
        out[name + ".seeds"] = ref.lsd_map(frame_of(name), min_norm)["sorted_rc"].astype(np.uint16)
        print(name, lines.shape, out[name + ".seeds"].shape)
    np.savez_compressed(os.path.join(HERE, "lsd_segments.npz"), **out)
