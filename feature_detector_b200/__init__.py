"""feature_detector_b200 -- B200-native dense feature-detection kernels behind the reference's interfaces.

Python here is plumbing only: a ctypes binding of the C ABI in ``include/fd_b200.h`` (implemented by the
sm_100a kernels in ``csrc/`` and built in-tree as ``libfd_b200.so``).  The binding takes numpy arrays on
the host side and raw device pointers (e.g. ``torch.Tensor.data_ptr()``) on the device side, so the same
calls serve the parity tests and ``bench.py``.

There is no CPU fallback: importing works anywhere (so the ABI can be inspected), but creating a
:class:`Context` without the built library or without a CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

__all__ = ["Context", "FdError", "HARRIS", "SHI_TOMAS", "FAST", "SAMPLE_BILINEAR", "SAMPLE_TRUNCATE", "library_path", "load_library",
           "DetectParams", "BriefParams", "LsdParams", "NnParams", "sparsify"]

HARRIS, SHI_TOMAS, FAST = 0, 1, 2
SAMPLE_BILINEAR, SAMPLE_TRUNCATE = 0, 1

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_STATUS = {0: "FD_OK", 1: "FD_ERR_INVALID_ARGUMENT", 2: "FD_ERR_NO_DEVICE", 3: "FD_ERR_CUDA", 4: "FD_ERR_OUT_OF_MEMORY",
           5: "FD_ERR_CAPACITY", 6: "FD_ERR_NOT_READY"}


class FdError(RuntimeError):
    def __init__(self, status, message=""):
        self.status = status
        super().__init__(f"{_STATUS.get(status, status)}: {message}")


class DetectParams(C.Structure):
    """fd_detect_params; defaults are the reference's (feature_point_detector.h:15-20 and the SubOptions)."""
    _fields_ = [("kind", C.c_int32), ("min_valid_response", C.c_float), ("min_feature_distance", C.c_int32),
                ("needed_feature_num", C.c_uint32), ("harris_alpha", C.c_float), ("fast_n", C.c_int32),
                ("fast_min_pixel_diff", C.c_int32), ("reserved", C.c_int32)]

    def __init__(self, kind=HARRIS, min_valid_response=0.1, min_feature_distance=15, needed_feature_num=200, harris_alpha=0.04,
                 fast_n=12, fast_min_pixel_diff=15):
        super().__init__(kind, min_valid_response, min_feature_distance, needed_feature_num, harris_alpha, fast_n, fast_min_pixel_diff, 0)


class BriefParams(C.Structure):
    _fields_ = [("length", C.c_int32), ("half_patch_size", C.c_int32), ("sampling", C.c_int32), ("reserved", C.c_int32)]

    def __init__(self, length=256, half_patch_size=8, sampling=SAMPLE_BILINEAR):
        super().__init__(length, half_patch_size, sampling, 0)


class LsdParams(C.Structure):
    _fields_ = [("min_valid_gradient_norm", C.c_float), ("want_sorted", C.c_int32)]

    def __init__(self, min_valid_gradient_norm=20.0, want_sorted=1):
        super().__init__(min_valid_gradient_norm, want_sorted)


class NnParams(C.Structure):
    """NNFeaturePointDetector::Options fields the post-processing reads (nn_feature_point_detector.h:22-31)."""
    _fields_ = [("min_response", C.c_float), ("invalid_boundary", C.c_int32), ("min_feature_distance", C.c_int32), ("max_features", C.c_uint32),
                ("reserved", C.c_int32)]

    def __init__(self, min_response=0.1, invalid_boundary=3, min_feature_distance=15, max_features=240):
        super().__init__(min_response, invalid_boundary, min_feature_distance, max_features, 0)


KEYPOINT_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("response", "<f4"), ("reserved", "<i4")])
CANDIDATE_DTYPE = np.dtype([("response", "<f4"), ("x", "<i4"), ("y", "<i4")])
MATCH_DTYPE = np.dtype([("train_index", "<i4"), ("distance", "<i4"), ("second_distance", "<i4"), ("reserved", "<i4")])


def library_path() -> str:
    return os.path.join(_HERE, "libfd_b200.so")


def load_library():
    """Load libfd_b200.so and declare its signatures.  Raises if it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise FdError(2, f"{path} is missing: build it with `make -C feature_detector_b200/csrc` (or __graft_entry__.build()); "
                         "there is no CPU fallback")
    L = C.CDLL(path)
    vp, i32p, u8p, f32p = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_uint8), C.POINTER(C.c_float)
    sig = {
        "fd_create": (C.c_int, [C.c_int, C.POINTER(vp)]),
        "fd_destroy": (C.c_int, [vp]),
        "fd_last_error": (C.c_char_p, [vp]),
        "fd_version": (C.c_char_p, []),
        "fd_set_stream": (C.c_int, [vp, vp]),
        "fd_own_stream": (vp, [vp]),
        "fd_sync": (C.c_int, [vp]),
        "fd_launch_count": (C.c_uint64, [vp]),
        "fd_upload_frames": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int]),
        "fd_bind_device_frames": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int]),
        "fd_set_existing_features": (C.c_int, [vp, f32p, i32p, C.c_int, C.c_int]),
        "fd_detect": (C.c_int, [vp, C.POINTER(DetectParams), C.c_int]),
        "fd_compute_candidates": (C.c_int, [vp, C.POINTER(DetectParams), C.c_int]),
        "fd_set_dense_outputs": (C.c_int, [vp, vp, vp]),
        "fd_download_keypoints": (C.c_int, [vp, vp, i32p, C.c_int]),
        "fd_download_candidates": (C.c_int, [vp, C.c_int, vp, C.c_int64, C.POINTER(C.c_int64)]),
        "fd_candidate_counts": (C.c_int, [vp, i32p]),
        "fd_device_keypoints": (C.c_int, [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_int)]),
        "fd_set_tile": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int]),
        "fd_device_candidates": (C.c_int, [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_uint32)]),
        "fd_export_candidates": (C.c_int, [vp, vp, C.c_int64, C.POINTER(C.c_int64)]),
        "fd_select_candidates": (C.c_int, [vp, C.POINTER(DetectParams), vp, vp, C.c_uint32, C.c_int, C.c_int, C.c_int]),
        "fd_sparsify": (C.c_int, [f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint8, C.c_uint8, u8p]),
        "fd_describe_selected": (C.c_int, [vp, C.POINTER(BriefParams)]),
        "fd_describe_points": (C.c_int, [vp, C.POINTER(BriefParams), f32p, i32p, C.c_int, C.c_int]),
        "fd_download_descriptors": (C.c_int, [vp, vp, C.c_int]),
        "fd_device_descriptors": (C.c_int, [vp, C.POINTER(vp), C.POINTER(C.c_int)]),
        "fd_lsd_field": (C.c_int, [vp, C.POINTER(LsdParams), vp, vp, vp, vp]),
        "fd_descriptors_as_float": (C.c_int, [vp, vp]),
        "fd_download_descriptors_float": (C.c_int, [vp, vp, C.c_int]),
        "fd_upload_floats": (C.c_int, [vp, C.c_int, C.POINTER(C.c_float), C.c_size_t, C.POINTER(vp)]),
        "fd_nn_select_from_heatmap": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.POINTER(NnParams), C.c_int]),
        "fd_nn_sample_descriptors": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, vp]),
        "fd_nn_download_descriptors": (C.c_int, [vp, vp, C.c_int]),
        "fd_nn_sample_descriptors_at": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int32), C.c_int, C.c_int,
                                                  C.POINTER(C.c_float)]),
        "fd_lsd_device_outputs": (C.c_int, [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]),
        "fd_lsd_download": (C.c_int, [vp, C.c_int, vp, vp, vp, C.c_int64, i32p]),
        "fd_host_alloc": (C.c_int, [C.POINTER(vp), C.c_size_t]),
        "fd_host_free": (C.c_int, [vp]),
        "fd_detect_describe_host": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.POINTER(DetectParams), C.POINTER(BriefParams), C.c_int, vp, i32p, vp, C.c_int]),
        "fd_match_consecutive": (C.c_int, [vp]),
        "fd_match_descriptors": (C.c_int, [vp, vp, vp, C.c_int, vp, vp, C.c_int, C.c_int, vp]),
        "fd_download_matches": (C.c_int, [vp, vp, C.c_int]),
        "fd_tiled_create": (C.c_int, [i32p, C.c_int, C.POINTER(vp)]),
        "fd_tiled_destroy": (C.c_int, [vp]),
        "fd_tiled_last_error": (C.c_char_p, [vp]),
        "fd_tiled_upload_frames": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int]),
        "fd_tiled_scatter_device_frames": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int]),
        "fd_tiled_tile_info": (C.c_int, [vp, C.c_int, i32p, i32p, i32p, C.POINTER(vp), C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(vp)]),
        "fd_tiled_exchange_halos": (C.c_int, [vp]),
        "fd_tiled_halo_bytes": (C.c_uint64, [vp]),
        "fd_tiled_compute_candidates": (C.c_int, [vp, C.POINTER(DetectParams), C.c_int]),
        "fd_tiled_detect": (C.c_int, [vp, C.POINTER(DetectParams), C.c_int]),
        "fd_tiled_sync": (C.c_int, [vp]),
        "fd_tiled_candidate_counts": (C.c_int, [vp, i32p]),
        "fd_tiled_device_candidates": (C.c_int, [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_uint32), i32p]),
        "fd_tiled_download_candidates": (C.c_int, [vp, C.c_int, vp, C.c_int64, C.POINTER(C.c_int64)]),
        "fd_tiled_download_keypoints": (C.c_int, [vp, vp, i32p, C.c_int]),
        "fd_tiled_root_context": (vp, [vp]),
        "fd_debug_fast_offset_bits": (C.c_int, [C.c_uint32, C.POINTER(C.c_uint32), i32p]),
        "fd_debug_run_length_lut": (C.c_int, [u8p]),
        "fd_debug_harris_trace_min": (C.c_int, [C.c_float, C.POINTER(C.c_float)]),
        "fd_debug_check_guards": (C.c_int, [vp, i32p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _LIB = L
    return L


def sparsify(features_xy, image_rows, image_cols, status_need_filter, status_after_filter, status=None, grid_rows=12, grid_cols=12):
    """FeaturePointDetector::SparsifyFeatures (feature_point_detector.cpp:27-52).  Returns the new status array."""
    L = load_library()
    f = np.ascontiguousarray(features_xy, np.float32).reshape(-1, 2)
    st = None if status is None else np.ascontiguousarray(status, np.uint8).copy()
    if st is None or len(st) != len(f):
        st = np.ones(len(f), np.uint8)  # :29-31
    rc = L.fd_sparsify(f.ctypes.data_as(C.POINTER(C.c_float)), len(f), image_rows, image_cols, grid_rows, grid_cols, status_need_filter,
                       status_after_filter, st.ctypes.data_as(C.POINTER(C.c_uint8)))
    if rc != 0:
        raise FdError(rc, "fd_sparsify")
    return st


class TiledDetector:
    """Large frames row-tiled over the GPUs of a box in ONE process (fd_tiled_*, csrc/fd_tiled.cu): tile k of every frame on
    devices[k] (ordinals may repeat), halo rows by device-to-device peer copies, candidate keys packed on devices[0] by a kernel
    that reads the peers' slots, selection there.  Results equal Context.detect on the whole frames."""

    def __init__(self, devices):
        self._lib = load_library()
        self._h = C.c_void_p()
        dev = np.ascontiguousarray(devices, np.int32)
        rc = self._lib.fd_tiled_create(dev.ctypes.data_as(C.POINTER(C.c_int32)), len(dev), C.byref(self._h))
        if rc != 0:
            raise FdError(rc, "fd_tiled_create failed (no CUDA device? there is no CPU fallback)")
        self.n_tiles = len(dev)
        self.rows = self.cols = self.n_frames = 0

    def _ck(self, rc):
        if rc != 0:
            raise FdError(rc, (self._lib.fd_tiled_last_error(self._h) or b"").decode())

    def close(self):
        if self._h:
            self._lib.fd_tiled_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def upload(self, frames: np.ndarray):
        f = np.ascontiguousarray(frames, np.uint8)
        if f.ndim == 2:
            f = f[None]
        self.n_frames, self.rows, self.cols = f.shape
        self._ck(self._lib.fd_tiled_upload_frames(self._h, f.ctypes.data_as(C.c_void_p), self.rows, self.cols, self.n_frames))

    def scatter_device(self, dev_ptr: int, rows: int, cols: int, n_frames: int = 1, pitch: int = 0, frame_stride: int = 0):
        pitch = pitch or cols
        self.n_frames, self.rows, self.cols = n_frames, rows, cols
        self._ck(self._lib.fd_tiled_scatter_device_frames(self._h, C.c_void_p(dev_ptr), rows, cols, pitch, frame_stride or pitch * rows, n_frames))

    def tile_info(self, tile: int):
        dev, first, count = C.c_int32(), C.c_int32(), C.c_int32()
        ptr, stream = C.c_void_p(), C.c_void_p()
        pitch, stride = C.c_int64(), C.c_int64()
        self._ck(self._lib.fd_tiled_tile_info(self._h, tile, C.byref(dev), C.byref(first), C.byref(count), C.byref(ptr), C.byref(pitch), C.byref(stride),
                                              C.byref(stream)))
        return {"device": dev.value, "own_first_row": first.value, "own_row_count": count.value, "ptr": ptr.value or 0, "pitch": pitch.value,
                "frame_stride": stride.value, "stream": stream.value or 0}

    def exchange_halos(self):
        self._ck(self._lib.fd_tiled_exchange_halos(self._h))

    @property
    def halo_bytes(self) -> int:
        return int(self._lib.fd_tiled_halo_bytes(self._h))

    def compute_candidates(self, params: DetectParams, cand_capacity_per_tile: int = 0):
        self._ck(self._lib.fd_tiled_compute_candidates(self._h, C.byref(params), cand_capacity_per_tile))

    def detect(self, params: DetectParams, cand_capacity_per_tile: int = 0):
        self._ck(self._lib.fd_tiled_detect(self._h, C.byref(params), cand_capacity_per_tile))

    def sync(self):
        self._ck(self._lib.fd_tiled_sync(self._h))

    def candidate_counts(self) -> np.ndarray:
        out = np.zeros(self.n_frames, np.int32)
        self._ck(self._lib.fd_tiled_candidate_counts(self._h, out.ctypes.data_as(C.POINTER(C.c_int32))))
        return out

    def candidates(self, frame: int = 0) -> np.ndarray:
        n = C.c_int64()
        self._ck(self._lib.fd_tiled_download_candidates(self._h, frame, None, 0, C.byref(n)))
        out = np.zeros(n.value, CANDIDATE_DTYPE)
        if n.value:
            self._ck(self._lib.fd_tiled_download_candidates(self._h, frame, out.ctypes.data_as(C.c_void_p), n.value, C.byref(n)))
        return out

    def keypoints(self, capacity: int):
        kp = np.zeros((self.n_frames, capacity), KEYPOINT_DTYPE)
        cnt = np.zeros(self.n_frames, np.int32)
        self._ck(self._lib.fd_tiled_download_keypoints(self._h, kp.ctypes.data_as(C.c_void_p), cnt.ctypes.data_as(C.POINTER(C.c_int32)), capacity))
        return kp, cnt


class Context:
    """One device + one stream (fd_context)."""

    def __init__(self, device: int = 0):
        self._lib = load_library()
        self._h = C.c_void_p()
        rc = self._lib.fd_create(device, C.byref(self._h))
        if rc != 0:
            raise FdError(rc, "fd_create failed (no CUDA device? there is no CPU fallback)")
        self.rows = self.cols = self.n_frames = 0
        self._keep = []

    # -- plumbing ------------------------------------------------------------------------------------
    def _ck(self, rc):
        if rc != 0:
            raise FdError(rc, (self._lib.fd_last_error(self._h) or b"").decode())

    def close(self):
        if self._h:
            self._lib.fd_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def set_stream(self, cuda_stream_handle):
        """Run on the given cudaStream_t handle (0 = the legacy default stream); None = back to the context's own stream."""
        if cuda_stream_handle is None:
            cuda_stream_handle = self._lib.fd_own_stream(self._h)
        self._ck(self._lib.fd_set_stream(self._h, C.c_void_p(cuda_stream_handle or 0)))

    def sync(self):
        self._ck(self._lib.fd_sync(self._h))

    @property
    def launch_count(self) -> int:
        return int(self._lib.fd_launch_count(self._h))

    def check_guards(self) -> int:
        """Verify the red zones round every context-owned buffer (contexts created under FD_B200_GUARD=1); returns how many were checked."""
        n = C.c_int32(0)
        self._ck(self._lib.fd_debug_check_guards(self._h, C.byref(n)))
        return n.value

    # -- frames --------------------------------------------------------------------------------------
    def upload(self, frames):
        """frames: (n, rows, cols) or (rows, cols) uint8 host array (numpy, or anything exposing a host data pointer)."""
        a = np.ascontiguousarray(frames, np.uint8)
        if a.ndim == 2:
            a = a[None]
        n, rows, cols = a.shape
        self._ck(self._lib.fd_upload_frames(self._h, a.ctypes.data_as(C.c_void_p), rows, cols, n))
        self.rows, self.cols, self.n_frames = rows, cols, n

    def upload_ptr(self, host_ptr: int, rows: int, cols: int, n_frames: int):
        """Upload from a raw host pointer (e.g. a pinned torch tensor's data_ptr())."""
        self._ck(self._lib.fd_upload_frames(self._h, C.c_void_p(host_ptr), rows, cols, n_frames))
        self.rows, self.cols, self.n_frames = rows, cols, n_frames

    def bind_device(self, dev_ptr: int, rows: int, cols: int, n_frames: int, pitch: int | None = None, frame_stride: int | None = None):
        pitch = cols if pitch is None else pitch
        frame_stride = pitch * rows if frame_stride is None else frame_stride
        self._ck(self._lib.fd_bind_device_frames(self._h, C.c_void_p(dev_ptr), rows, cols, pitch, frame_stride, n_frames))
        self.rows, self.cols, self.n_frames = rows, cols, n_frames

    def set_existing_features(self, per_frame_xy):
        """per_frame_xy: list (one entry per frame) of (k, 2) float arrays; empty list clears."""
        if not per_frame_xy:
            self._ck(self._lib.fd_set_existing_features(self._h, None, None, 0, 0))
            return
        cap = max(1, max(len(np.asarray(p).reshape(-1, 2)) for p in per_frame_xy))
        xy = np.zeros((len(per_frame_xy), cap, 2), np.float32)
        counts = np.zeros(len(per_frame_xy), np.int32)
        for f, p in enumerate(per_frame_xy):
            p = np.asarray(p, np.float32).reshape(-1, 2)
            xy[f, :len(p)] = p
            counts[f] = len(p)
        self._ck(self._lib.fd_set_existing_features(self._h, xy.ctypes.data_as(C.POINTER(C.c_float)), counts.ctypes.data_as(C.POINTER(C.c_int32)),
                                                    cap, len(per_frame_xy)))

    # -- detection -----------------------------------------------------------------------------------
    def set_dense_outputs(self, dev_response_ptr: int = 0, dev_score_ptr: int = 0):
        self._ck(self._lib.fd_set_dense_outputs(self._h, C.c_void_p(dev_response_ptr or 0), C.c_void_p(dev_score_ptr or 0)))

    def detect(self, params: DetectParams, cand_capacity: int = 0):
        self._ck(self._lib.fd_detect(self._h, C.byref(params), cand_capacity))

    def compute_candidates(self, params: DetectParams, cand_capacity: int = 0):
        self._ck(self._lib.fd_compute_candidates(self._h, C.byref(params), cand_capacity))

    def keypoints(self, kp_capacity: int):
        """Returns (kp structured array (n_frames, kp_capacity), counts (n_frames,))."""
        kp = np.zeros((self.n_frames, kp_capacity), KEYPOINT_DTYPE)
        counts = np.zeros(self.n_frames, np.int32)
        self._ck(self._lib.fd_download_keypoints(self._h, kp.ctypes.data_as(C.c_void_p), counts.ctypes.data_as(C.POINTER(C.c_int32)), kp_capacity))
        return kp, counts

    def keypoints_into(self, kp: np.ndarray, counts: np.ndarray):
        """Like keypoints(), into caller-owned arrays (e.g. views of pinned memory): kp (n_frames, capacity) KEYPOINT_DTYPE, counts int32."""
        assert kp.dtype == KEYPOINT_DTYPE and kp.flags.c_contiguous and kp.shape[0] == self.n_frames and counts.dtype == np.int32
        self._ck(self._lib.fd_download_keypoints(self._h, kp.ctypes.data_as(C.c_void_p), counts.ctypes.data_as(C.POINTER(C.c_int32)), kp.shape[1]))

    def keypoint_counts(self):
        counts = np.zeros(self.n_frames, np.int32)
        self._ck(self._lib.fd_download_keypoints(self._h, None, counts.ctypes.data_as(C.POINTER(C.c_int32)), 0))
        return counts

    def candidate_counts(self):
        counts = np.zeros(self.n_frames, np.int32)
        self._ck(self._lib.fd_candidate_counts(self._h, counts.ctypes.data_as(C.POINTER(C.c_int32))))
        return counts

    def candidates(self, frame: int = 0):
        n = C.c_int64(0)
        self._ck(self._lib.fd_download_candidates(self._h, frame, None, 0, C.byref(n)))
        out = np.zeros(n.value, CANDIDATE_DTYPE)
        self._ck(self._lib.fd_download_candidates(self._h, frame, out.ctypes.data_as(C.c_void_p), n.value, C.byref(n)))
        return out

    # -- row tiles of one large frame ------------------------------------------------------------------
    def set_tile(self, row_offset: int, own_first_row: int, own_row_count: int, full_rows: int):
        """The bound frames are rows [row_offset, row_offset + rows) of an image full_rows tall; full_rows <= 0 clears."""
        self._ck(self._lib.fd_set_tile(self._h, row_offset, own_first_row, own_row_count, full_rows))

    def device_candidates(self):
        """(dev_keys ptr, dev_counts ptr, capacity) of the last fd_compute_candidates."""
        keys, cnt, cap = C.c_void_p(), C.c_void_p(), C.c_uint32(0)
        self._ck(self._lib.fd_device_candidates(self._h, C.byref(keys), C.byref(cnt), C.byref(cap)))
        return keys.value, cnt.value, cap.value

    def export_candidates(self, dev_dst: int, dst_capacity: int):
        """Pack all frames' candidate keys into device memory at dev_dst (int64 / uint64 elements); returns per-frame counts."""
        counts = np.zeros(self.n_frames, np.int64)
        self._ck(self._lib.fd_export_candidates(self._h, C.c_void_p(dev_dst), dst_capacity, counts.ctypes.data_as(C.POINTER(C.c_int64))))
        return counts

    def select_candidates(self, params: DetectParams, dev_keys: int, dev_counts: int, capacity: int, rows: int, cols: int, n_frames: int = 1):
        self._ck(self._lib.fd_select_candidates(self._h, C.byref(params), C.c_void_p(dev_keys), C.c_void_p(dev_counts), capacity, rows, cols, n_frames))
        self.n_frames = n_frames

    def device_keypoints(self):
        kp, cnt, cap = C.c_void_p(), C.c_void_p(), C.c_int(0)
        self._ck(self._lib.fd_device_keypoints(self._h, C.byref(kp), C.byref(cnt), C.byref(cap)))
        return kp.value, cnt.value, cap.value

    # -- BRIEF ---------------------------------------------------------------------------------------
    def describe_selected(self, params: BriefParams):
        self._ck(self._lib.fd_describe_selected(self._h, C.byref(params)))

    def describe_points(self, params: BriefParams, per_frame_xy):
        cap = max(1, max(len(np.asarray(p).reshape(-1, 2)) for p in per_frame_xy))
        xy = np.zeros((len(per_frame_xy), cap, 2), np.float32)
        counts = np.zeros(len(per_frame_xy), np.int32)
        for f, p in enumerate(per_frame_xy):
            p = np.asarray(p, np.float32).reshape(-1, 2)
            xy[f, :len(p)] = p
            counts[f] = len(p)
        self._ck(self._lib.fd_describe_points(self._h, C.byref(params), xy.ctypes.data_as(C.POINTER(C.c_float)),
                                              counts.ctypes.data_as(C.POINTER(C.c_int32)), cap, len(per_frame_xy)))
        return cap

    def descriptors(self, kp_capacity: int):
        """(n_frames, kp_capacity, 32) uint8; bit i of a descriptor is (byte i//8 >> i%8) & 1."""
        d = np.zeros((self.n_frames, kp_capacity, 32), np.uint8)
        self._ck(self._lib.fd_download_descriptors(self._h, d.ctypes.data_as(C.c_void_p), kp_capacity))
        return d

    def descriptors_float(self, kp_capacity: int, length: int = 256):
        """(n_frames, kp_capacity, length) float32 of +1 / -1: the std::vector<Vec> overload of Descriptor::Compute (descriptor.h:43-62)."""
        self._ck(self._lib.fd_descriptors_as_float(self._h, None))
        d = np.zeros((self.n_frames, kp_capacity, length), np.float32)
        self._ck(self._lib.fd_download_descriptors_float(self._h, d.ctypes.data_as(C.c_void_p), kp_capacity))
        return d

    def descriptors_into(self, desc: np.ndarray):
        """Like descriptors(), into a caller-owned (n_frames, capacity, 32) uint8 array."""
        assert desc.dtype == np.uint8 and desc.flags.c_contiguous and desc.shape[0] == self.n_frames and desc.shape[2] == 32
        self._ck(self._lib.fd_download_descriptors(self._h, desc.ctypes.data_as(C.c_void_p), desc.shape[1]))

    # -- NN detector post-processing ---------------------------------------------------------------------
    def detect_describe_host(self, frames, detect: DetectParams, brief: BriefParams | None, kp_capacity: int, cand_capacity: int = 0):
        """One host-to-host call (fd_detect_describe_host): returns (keypoints (n, cap) KEYPOINT_DTYPE, counts (n,), descriptors (n, cap, 32) or None)."""
        a = np.ascontiguousarray(frames, np.uint8)
        if a.ndim == 2:
            a = a[None]
        n, rows, cols = a.shape
        kp = np.zeros((n, kp_capacity), KEYPOINT_DTYPE)
        cnt = np.zeros(n, np.int32)
        desc = np.zeros((n, kp_capacity, 32), np.uint8) if brief is not None else None
        self._ck(self._lib.fd_detect_describe_host(self._h, a.ctypes.data_as(C.c_void_p), rows, cols, n, C.byref(detect), C.byref(brief) if brief is not None else None,
                                                   cand_capacity, kp.ctypes.data_as(C.c_void_p), cnt.ctypes.data_as(C.POINTER(C.c_int32)),
                                                   desc.ctypes.data_as(C.c_void_p) if desc is not None else None, kp_capacity))
        self.rows, self.cols, self.n_frames = rows, cols, n
        return kp, cnt, desc

    # -- Hamming matching of the packed descriptors (no reference counterpart) ----------------------------
    def match_selected(self):
        """Frame f of the last described set against frame f + 1 (fd_match_consecutive); results stay on the device."""
        self._ck(self._lib.fd_match_consecutive(self._h))

    def matches(self, kp_capacity: int) -> np.ndarray:
        """(n_frames - 1, kp_capacity) MATCH_DTYPE of the last match_selected()."""
        out = np.zeros((max(self.n_frames - 1, 0), kp_capacity), MATCH_DTYPE)
        buf = out if out.size else np.zeros((1, kp_capacity), MATCH_DTYPE)
        self._ck(self._lib.fd_download_matches(self._h, buf.ctypes.data_as(C.c_void_p), kp_capacity))
        return out

    def match_descriptors(self, dev_desc_a: int, dev_counts_a: int, capacity_a: int, dev_desc_b: int, dev_counts_b: int, capacity_b: int, n_pairs: int,
                          dev_out: int):
        self._ck(self._lib.fd_match_descriptors(self._h, C.c_void_p(dev_desc_a), C.c_void_p(dev_counts_a), capacity_a, C.c_void_p(dev_desc_b),
                                                C.c_void_p(dev_counts_b), capacity_b, n_pairs, C.c_void_p(dev_out)))

    def nn_select(self, dev_heatmap: int, rows: int, cols: int, n_frames: int, params: NnParams, cand_capacity: int = 0):
        """Heat maps (device pointer, n_frames x rows x cols float32) -> keypoints (fetch with keypoints())."""
        self._ck(self._lib.fd_nn_select_from_heatmap(self._h, C.c_void_p(dev_heatmap), rows, cols, n_frames, C.byref(params), cand_capacity))
        self.rows, self.cols, self.n_frames = rows, cols, n_frames

    def nn_sample_descriptors(self, dev_maps: int, channels: int, map_rows: int, map_cols: int, dev_out: int = 0):
        self._ck(self._lib.fd_nn_sample_descriptors(self._h, C.c_void_p(dev_maps), channels, map_rows, map_cols, C.c_void_p(dev_out or 0)))
        self._nn_channels = channels

    def nn_descriptors(self, kp_capacity: int):
        """(n_frames, kp_capacity, channels) float32 from the context-owned buffer."""
        d = np.zeros((self.n_frames, kp_capacity, self._nn_channels), np.float32)
        self._ck(self._lib.fd_nn_download_descriptors(self._h, d.ctypes.data_as(C.c_void_p), kp_capacity))
        return d

    def nn_descriptors_at(self, dev_maps: int, channels: int, map_rows: int, map_cols: int, per_frame_xy):
        """Descriptors at caller-supplied points: list (one entry per frame) of (k, 2) float arrays -> list of (k, channels) float32."""
        cap = max(1, max(len(np.asarray(p).reshape(-1, 2)) for p in per_frame_xy))
        xy = np.zeros((len(per_frame_xy), cap, 2), np.float32)
        counts = np.zeros(len(per_frame_xy), np.int32)
        for f, p in enumerate(per_frame_xy):
            p = np.asarray(p, np.float32).reshape(-1, 2)
            xy[f, :len(p)] = p
            counts[f] = len(p)
        out = np.zeros((len(per_frame_xy), cap, channels), np.float32)
        self._ck(self._lib.fd_nn_sample_descriptors_at(self._h, C.c_void_p(dev_maps), channels, map_rows, map_cols, xy.ctypes.data_as(C.POINTER(C.c_float)),
                                                       counts.ctypes.data_as(C.POINTER(C.c_int32)), cap, len(per_frame_xy),
                                                       out.ctypes.data_as(C.POINTER(C.c_float))))
        return [out[f, :counts[f]] for f in range(len(per_frame_xy))]

    # -- LSD -----------------------------------------------------------------------------------------
    def lsd_field(self, params: LsdParams, dev_norm: int = 0, dev_angle: int = 0, dev_sorted: int = 0, dev_n_valid: int = 0):
        self._ck(self._lib.fd_lsd_field(self._h, C.byref(params), C.c_void_p(dev_norm or 0), C.c_void_p(dev_angle or 0), C.c_void_p(dev_sorted or 0),
                                        C.c_void_p(dev_n_valid or 0)))

    def lsd_download(self, frame: int = 0, want_sorted: bool = True):
        """Returns dict(norm, angle (rows, cols) float32 maps with zero last row/column, sorted_idx, n_valid)."""
        norm = np.zeros((self.rows, self.cols), np.float32)
        angle = np.zeros((self.rows, self.cols), np.float32)
        cap = self.rows * self.cols if want_sorted else 0
        sorted_idx = np.zeros(cap, np.int32)
        n = C.c_int32(0)
        self._ck(self._lib.fd_lsd_download(self._h, frame, norm.ctypes.data_as(C.c_void_p), angle.ctypes.data_as(C.c_void_p),
                                           sorted_idx.ctypes.data_as(C.c_void_p) if want_sorted else None, cap, C.byref(n)))
        return {"norm": norm, "angle": angle, "sorted_idx": sorted_idx[:n.value].copy(), "n_valid": n.value}


def unpack_bits(desc: np.ndarray, length: int = 256) -> np.ndarray:
    """(…, 32) packed descriptors -> (…, length) array of 0/1 (LSB-first within each byte)."""
    return np.unpackbits(desc, axis=-1, bitorder="little")[..., :length]
