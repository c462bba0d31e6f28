"""TEST INFRASTRUCTURE — ctypes access to the two CPU checkers.

* ``Ref``  : oracle/_ref/libfd_ref.so — the UNMODIFIED reference sources compiled in place
             (oracle/Makefile).  Present wherever the repo was built with /root/reference mounted
             (it travels to the GPU box as a prebuilt file).
* ``Port`` : oracle/libfd_oracle.so — the plain-C restatement in oracle/fd_oracle.c.

Only tests/, ``__graft_entry__.smoke()`` and bench.py's ``cpu_baseline`` / ``--impl reference``
legs may import this module.  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(_HERE, "_ref", "libfd_ref.so")
PORT_SO = os.path.join(_HERE, "libfd_oracle.so")

HARRIS, SHI_TOMAS, FAST = 0, 1, 2

_u8p = C.POINTER(C.c_uint8)
_f32p = C.POINTER(C.c_float)
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_i16p = C.POINTER(C.c_int16)


def _p(a, t):
    return None if a is None else a.ctypes.data_as(t)


def _img(img):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    assert img.ndim == 2
    return img


def fnv1a64(data: bytes) -> int:
    """64-bit FNV-1a, the hash SURVEY.md section 8c quotes its known answers in."""
    h = 14695981039346656037
    for chunk in (data,):
        for b in chunk:
            h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def fnv1a64_np(arr: np.ndarray) -> int:
    """Vectorised-ish FNV-1a over the little-endian bytes of ``arr`` (pure Python loop over bytes is
    too slow for MB-sized maps; this one runs the recurrence in C via int.from_bytes chunks)."""
    data = np.ascontiguousarray(arr).tobytes()
    h = 14695981039346656037
    prime = 1099511628211
    mask = 0xFFFFFFFFFFFFFFFF
    for b in data:
        h = ((h ^ b) * prime) & mask
    return h


class _Detectors:
    """Shared call surface of Ref and Port (same C signatures by construction)."""

    def __init__(self, path, prefix):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        self.prefix = prefix
        L = self.lib
        f = getattr(L, prefix + "detect")
        f.restype = C.c_int
        f.argtypes = [C.c_int, _u8p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_uint32, C.c_int, _f32p, C.c_int, C.c_int,
                      C.POINTER(C.c_int), _f32p, _i32p, C.c_int64, _i64p, _f32p, _i32p]
        f = getattr(L, prefix + "fast_score_map")
        f.restype = None
        f.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_int, _u8p]
        f = getattr(L, prefix + "brief")
        f.restype = C.c_int
        f.argtypes = [_u8p, C.c_int, C.c_int, _f32p, C.c_int, C.c_int, C.c_int, _u8p]
        f = getattr(L, prefix + "lsd_map")
        f.restype = C.c_int
        f.argtypes = [_u8p, C.c_int, C.c_int, C.c_float, _f32p, _f32p, _u8p, _i32p, C.c_int64, _i64p]
        f = getattr(L, prefix + "sparsify")
        f.restype = None
        f.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint8, C.c_uint8, _u8p, C.c_int]

    # -- point detectors -------------------------------------------------------------------------
    def detect(self, kind, img, min_response, min_distance, needed, fast_n=0, pre=None, want_response=False,
               want_mask=False, want_candidates=True):
        """Returns dict(ok, features (n,2) f32, cand_resp, cand_xy, response, mask)."""
        img = _img(img)
        rows, cols = img.shape
        pre = np.zeros((0, 2), np.float32) if pre is None else np.ascontiguousarray(pre, np.float32).reshape(-1, 2)
        max_feats = int(needed) + len(pre) + 8
        feats = np.zeros((max_feats, 2), np.float32)
        feats[:len(pre)] = pre
        n_out = C.c_int(0)
        n_cand = C.c_int64(0)
        max_cand = rows * cols if want_candidates else 0
        cand_resp = np.zeros(max_cand, np.float32) if want_candidates else None
        cand_xy = np.zeros((max_cand, 2), np.int32) if want_candidates else None
        resp = np.zeros((rows, cols), np.float32) if want_response else None
        mask = np.zeros((rows, cols), np.int32) if want_mask else None
        rc = getattr(self.lib, self.prefix + "detect")(
            kind, _p(img, _u8p), rows, cols, float(min_response), int(min_distance), int(needed), int(fast_n),
            _p(feats, _f32p), len(pre), max_feats, C.byref(n_out), _p(cand_resp, _f32p), _p(cand_xy, _i32p), max_cand,
            C.byref(n_cand), _p(resp, _f32p), _p(mask, _i32p))
        if rc < 0:
            raise RuntimeError(f"{self.prefix}detect buffer too small ({rc})")
        out = {"ok": bool(rc), "features": feats[:n_out.value].copy(), "n_cand": n_cand.value, "response": resp, "mask": mask}
        if want_candidates:
            out["cand_resp"] = cand_resp[:n_cand.value].copy()
            out["cand_xy"] = cand_xy[:n_cand.value].copy()
        return out

    def fast_score_map(self, img, fast_n=12, diff=15):
        img = _img(img)
        out = np.zeros(img.shape, np.uint8)
        getattr(self.lib, self.prefix + "fast_score_map")(_p(img, _u8p), img.shape[0], img.shape[1], fast_n, diff, _p(out, _u8p))
        return out

    # -- BRIEF ----------------------------------------------------------------------------------
    def brief(self, img, kp_xy, length=256, half_patch=8):
        """Returns (ok, bits (n, length) uint8 of 0/1)."""
        img = _img(img)
        kp = np.ascontiguousarray(kp_xy, np.float32).reshape(-1, 2)
        bits = np.zeros((len(kp), length), np.uint8)
        ok = getattr(self.lib, self.prefix + "brief")(_p(img, _u8p), img.shape[0], img.shape[1], _p(kp, _f32p), len(kp), length,
                                                     half_patch, _p(bits, _u8p))
        return bool(ok), bits

    def brief_vec(self, img, kp_xy, length=256, half_patch=8):
        """The std::vector<Vec> overload: returns (ok, (n, length) float32 of +1 / -1)."""
        img = _img(img)
        kp = np.ascontiguousarray(kp_xy, np.float32).reshape(-1, 2)
        out = np.zeros((len(kp), length), np.float32)
        fn = getattr(self.lib, self.prefix + "brief_vec")
        fn.restype = C.c_int
        fn.argtypes = [_u8p, C.c_int, C.c_int, _f32p, C.c_int, C.c_int, C.c_int, _f32p]
        ok = fn(_p(img, _u8p), img.shape[0], img.shape[1], _p(kp, _f32p), len(kp), length, half_patch, _p(out, _f32p))
        return bool(ok), out

    # -- LSD map --------------------------------------------------------------------------------
    def lsd_map(self, img, min_norm=20.0):
        """Returns dict(norm, angle, valid ((rows-1),(cols-1)), sorted_rc (n,2))."""
        img = _img(img)
        rows, cols = img.shape
        shp = (rows - 1, cols - 1)
        norm = np.zeros(shp, np.float32)
        angle = np.zeros(shp, np.float32)
        valid = np.zeros(shp, np.uint8)
        cap = shp[0] * shp[1]
        sorted_rc = np.zeros((cap, 2), np.int32)
        n = C.c_int64(0)
        rc = getattr(self.lib, self.prefix + "lsd_map")(_p(img, _u8p), rows, cols, float(min_norm), _p(norm, _f32p), _p(angle, _f32p),
                                                       _p(valid, _u8p), _p(sorted_rc, _i32p), cap, C.byref(n))
        if rc != 1:
            raise RuntimeError(f"{self.prefix}lsd_map failed ({rc})")
        return {"norm": norm, "angle": angle, "valid": valid, "sorted_rc": sorted_rc[:n.value].copy()}

    def nn_select(self, heatmap, min_response=0.1, invalid_boundary=3, min_distance=15, max_features=240, pre=None):
        """NN detector post-processing on a heat map: returns dict(features (n,2) f32 incl. the pre-existing ones, n_cand)."""
        hm = np.ascontiguousarray(heatmap, np.float32)
        rows, cols = hm.shape
        pre = np.zeros((0, 2), np.float32) if pre is None else np.ascontiguousarray(pre, np.float32).reshape(-1, 2)
        max_feats = int(max_features) + len(pre) + 8
        feats = np.zeros((max_feats, 2), np.float32)
        feats[:len(pre)] = pre
        n_out, n_cand = C.c_int(0), C.c_int64(0)
        fn = getattr(self.lib, self.prefix + "nn_select")
        fn.restype = C.c_int
        fn.argtypes = [_f32p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, _f32p, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int64)]
        if fn(_p(hm, _f32p), rows, cols, float(min_response), int(invalid_boundary), int(min_distance), int(max_features), _p(feats, _f32p), len(pre),
              max_feats, C.byref(n_out), C.byref(n_cand)) != 1:
            raise RuntimeError(self.prefix + "nn_select failed")
        return {"features": feats[:n_out.value].copy(), "n_cand": n_cand.value}

    def nn_descriptors(self, feats_xy, maps):
        """maps: (channels, map_rows, map_cols) f32; returns (n, channels) f32."""
        f = np.ascontiguousarray(feats_xy, np.float32).reshape(-1, 2)
        m = np.ascontiguousarray(maps, np.float32)
        ch, mr, mc = m.shape
        out = np.zeros((len(f), ch), np.float32)
        fn = getattr(self.lib, self.prefix + "nn_descriptors")
        fn.restype = C.c_int
        fn.argtypes = [_f32p, C.c_int, _f32p, C.c_int, C.c_int, C.c_int, _f32p]
        if fn(_p(f, _f32p), len(f), _p(m, _f32p), ch, mr, mc, _p(out, _f32p)) != 1:
            raise RuntimeError(self.prefix + "nn_descriptors failed")
        return out

    def sparsify(self, feats_xy, rows, cols, need, after, status, grid_rows=12, grid_cols=12):
        f = np.ascontiguousarray(feats_xy, np.float32).reshape(-1, 2)
        st = np.ascontiguousarray(status, np.uint8).copy()
        if len(st) != len(f):
            st = np.ones(len(f), np.uint8)  # feature_point_detector.cpp:29-31
        getattr(self.lib, self.prefix + "sparsify")(_p(f, _f32p), len(f), rows, cols, grid_rows, grid_cols, need, after, _p(st, _u8p), len(st))
        return st


class Ref(_Detectors):
    """The reference itself (compiled in place)."""

    def __init__(self, path=REF_SO):
        super().__init__(path, "ref_")
        L = self.lib
        L.ref_brief_pattern.restype = None
        L.ref_brief_pattern.argtypes = [_i16p]
        L.ref_lsd_detect.restype = C.c_int
        L.ref_lsd_detect.argtypes = [_u8p, C.c_int, C.c_int, C.c_uint32, C.c_float, _f32p, C.c_int, C.POINTER(C.c_int)]
        L.ref_bench_points.restype = C.c_double
        L.ref_bench_points.argtypes = [C.c_int, _u8p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_uint32, C.c_int, C.c_int,
                                       C.c_int, C.c_int, _i64p]
        L.ref_bench_lsd.restype = C.c_double
        L.ref_bench_lsd.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, _i64p]

    def brief_pattern(self):
        out = np.zeros(1024, np.int16)
        self.lib.ref_brief_pattern(_p(out, _i16p))
        return out.reshape(256, 4)

    def lsd_detect(self, img, needed=200, min_norm=20.0, max_lines=20000):
        img = _img(img)
        lines = np.zeros((max_lines, 4), np.float32)
        n = C.c_int(0)
        rc = self.lib.ref_lsd_detect(_p(img, _u8p), img.shape[0], img.shape[1], needed, float(min_norm), _p(lines, _f32p), max_lines, C.byref(n))
        if rc < 0:
            raise RuntimeError("ref_lsd_detect: too many lines")
        return bool(rc), lines[:n.value].copy()

    def bench_points(self, kind, frames, min_response, min_distance, needed, fast_n=0, brief_length=0, brief_half_patch=8, n_threads=1):
        """Seconds of wall clock for DetectGoodFeatures (+BRIEF) over ``frames`` (n, rows, cols) u8, and totals."""
        frames = np.ascontiguousarray(frames, np.uint8)
        n, rows, cols = frames.shape
        totals = np.zeros(3, np.int64)
        sec = self.lib.ref_bench_points(kind, _p(frames, _u8p), n, rows, cols, float(min_response), int(min_distance), int(needed),
                                        int(fast_n), int(brief_length), int(brief_half_patch), int(n_threads), _p(totals, _i64p))
        return sec, totals

    def bench_lsd(self, frames, min_norm=20.0, full_detect=False, n_threads=1):
        frames = np.ascontiguousarray(frames, np.uint8)
        n, rows, cols = frames.shape
        totals = np.zeros(2, np.int64)
        sec = self.lib.ref_bench_lsd(_p(frames, _u8p), n, rows, cols, float(min_norm), int(full_detect), int(n_threads), _p(totals, _i64p))
        return sec, totals


class Port(_Detectors):
    """The plain-C restatement (oracle/fd_oracle.c)."""

    def __init__(self, path=PORT_SO):
        super().__init__(path, "orc_")
        L = self.lib
        L.orc_bench_points.restype = C.c_double
        L.orc_bench_points.argtypes = [C.c_int, _u8p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_uint32, C.c_int, C.c_int,
                                       C.c_int, C.c_int, _i64p]

    def bench_points(self, kind, frames, min_response, min_distance, needed, fast_n=0, brief_length=0, brief_half_patch=8, n_threads=1):
        frames = np.ascontiguousarray(frames, np.uint8)
        n, rows, cols = frames.shape
        totals = np.zeros(3, np.int64)
        sec = self.lib.orc_bench_points(kind, _p(frames, _u8p), n, rows, cols, float(min_response), int(min_distance), int(needed),
                                        int(fast_n), int(brief_length), int(brief_half_patch), int(n_threads), _p(totals, _i64p))
        return sec, totals


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def have_port() -> bool:
    return os.path.exists(PORT_SO)


def best_checker():
    """The reference when its build is present, else the port."""
    return Ref() if have_ref() else Port()
