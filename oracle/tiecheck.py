"""TEST INFRASTRUCTURE -- tie-aware comparison of two greedy selections.

The reference sorts candidates with an unstable std::sort (feature_point_detector.cpp:58), so candidates
with EQUAL response may be visited in any order; this framework fixes raster order.  Two results are
"identical except for ties" when
  (1) both sorted candidate lists carry the same response sequence and, inside every run of equal
      responses, the same set of pixels; and
  (2) replaying the reference's greedy walk (feature_point_detector.cpp:62-71) over each list reproduces
      the feature list that came with it.
"""
from __future__ import annotations

import numpy as np


def greedy_replay(cand_xy, rows, cols, min_distance, needed, pre=None):
    """Pure-Python restatement of SelectGoodFeatures over an already sorted candidate list."""
    mask = np.ones((rows, cols), bool)
    feats = [] if pre is None else [tuple(map(float, p)) for p in np.asarray(pre).reshape(-1, 2)]
    d = int(min_distance)
    for (x, y) in feats:
        r, c = int(y), int(x)
        mask[max(r - d, 0):r + d + 1, max(c - d, 0):c + d + 1] = False
    n_pre = len(feats)
    for x, y in np.asarray(cand_xy).reshape(-1, 2):
        if mask[y, x]:
            feats.append((float(x), float(y)))
            if len(feats) >= needed:
                break
            mask[max(y - d, 0):y + d + 1, max(x - d, 0):x + d + 1] = False
    return np.array(feats, np.float32).reshape(-1, 2), n_pre


def same_up_to_ties(resp_a, xy_a, resp_b, xy_b) -> bool:
    resp_a = np.asarray(resp_a, np.float32)
    resp_b = np.asarray(resp_b, np.float32)
    if resp_a.shape != resp_b.shape or not np.array_equal(resp_a.view(np.uint32), resp_b.view(np.uint32)):
        return False
    if len(resp_a) and np.any(np.diff(resp_a) > 0):
        return False  # not sorted descending
    xy_a = np.asarray(xy_a).reshape(-1, 2)
    xy_b = np.asarray(xy_b).reshape(-1, 2)
    # inside each run of equal responses compare as sets: sort both by (response desc, y, x)
    ka = np.lexsort((xy_a[:, 0], xy_a[:, 1], -resp_a.astype(np.float64)))
    kb = np.lexsort((xy_b[:, 0], xy_b[:, 1], -resp_b.astype(np.float64)))
    return bool(np.array_equal(xy_a[ka], xy_b[kb]))
