"""CPU suite, part 2: the C-ABI library loads, exports every symbol include/fd_b200.h declares, fails
loudly without a GPU, and its host-built tables are right."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import feature_detector_b200 as fd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "fd_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fd_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = fd.load_library()
    names = _declared_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/fd_b200.h but not exported"


def test_python_binding_covers_the_header():
    lib = fd.load_library()
    for n in _declared_functions():
        assert getattr(lib, n).argtypes is not None, f"binding for {n} missing"


def test_no_cpu_fallback_without_device():
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(fd.FdError) as e:
        fd.Context(0)
    assert e.value.status == 2  # FD_ERR_NO_DEVICE


def test_product_does_not_touch_the_oracle():
    """No file of the product package may import, link or open anything under oracle/."""
    pkg = os.path.join(ROOT, "feature_detector_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in text.replace("oracle/ port", ""), f"{f} mentions the oracle"
                assert "libfd_ref" not in text and "libfd_oracle" not in text


def test_fast_offset_table_matches_the_float_loop():
    lib = fd.load_library()
    n = 300_000
    out = np.zeros(n, np.uint32)
    nseg = C.c_int32(0)
    assert lib.fd_debug_fast_offset_bits(n, out.ctypes.data_as(C.POINTER(C.c_uint32)), C.byref(nseg)) == 0
    off = np.float32(1e-5)
    inc = np.float32(1e-5)
    exp = np.empty(n, np.float32)
    for k in range(n):  # feature_point_fast_detector.cpp:85,93
        exp[k] = off
        off = np.float32(off + inc)
    assert np.array_equal(exp.view(np.uint32), out)
    assert 8 <= nseg.value <= 64
    # the offset crosses 0.1 near the 10 000th masked-in pixel (rounding drift moves it a little)
    first = int(np.argmax(out.view(np.float32) > 0.1))
    assert 9990 <= first <= 10010


def test_fast_run_length_lut():
    lib = fd.load_library()
    lut = np.zeros(65536, np.uint8)
    assert lib.fd_debug_run_length_lut(lut.ctypes.data_as(C.POINTER(C.c_uint8))) == 0
    rng = np.random.default_rng(0)
    for m in [0, 0xFFFF, 0x8001, 0x00FF, 0xF00F, 0x7FFF, 0xAAAA] + rng.integers(0, 65536, 300).tolist():
        bits = [(m >> i) & 1 for i in range(16)]
        best = run = 0
        for i in range(32):  # walk the ring twice (fast.cpp:55-78)
            run = run + 1 if bits[i % 16] else 0
            best = max(best, run)
        assert lut[m] == min(best, 16)


def test_sparsify_host_logic(port, vectors):
    st = fd.sparsify(vectors["sparsify.features"], 480, 752, 1, 2, vectors["sparsify.status_in"])
    assert np.array_equal(st, vectors["sparsify.status_out"])
    rng = np.random.default_rng(5)
    for rows, cols, gr, gc in [(480, 752, 12, 12), (720, 1280, 8, 10), (100, 100, 3, 4), (5, 5, 12, 12), (480, 9, 12, 12)]:   # the last two: a step of 0
        f = np.stack([rng.uniform(-30, cols + 30, 300), rng.uniform(-30, rows + 30, 300)], 1).astype(np.float32)
        s0 = rng.integers(0, 3, 300).astype(np.uint8)
        assert np.array_equal(fd.sparsify(f, rows, cols, 1, 2, s0, gr, gc), port.sparsify(f, rows, cols, 1, 2, s0, gr, gc))
    # size mismatch -> all ones first (feature_point_detector.cpp:29-31)
    assert np.array_equal(fd.sparsify(f, 480, 752, 1, 2, None), port.sparsify(f, 480, 752, 1, 2, np.zeros(0, np.uint8)))


def test_sparse_fast_filter_is_conservative():
    """The facts the sparse FAST kernel's phase A rests on (fd_fast_sparse.cu header), checked exhaustively on the CPU:
    every ring mask whose longest circular run is >= 4 contains a compass position (ring 0, 4, 8, 12), every one with a run >= 8
    contains a vertical AND a horizontal compass position, every one with a run >= 12 passes the pre-check's closed form or
    contains 3 compass positions; and the packed-byte threshold test `(|a - b| & mask) != 0` with mask = bits at or above the largest
    power of two <= diff + 1 never misses a pixel with |a - b| > diff (exact when diff + 1 is a power of two, e.g. the default 15)."""
    lib = fd.load_library()
    lut = np.zeros(65536, np.uint8)
    assert lib.fd_debug_run_length_lut(lut.ctypes.data_as(C.POINTER(C.c_uint8))) == 0
    m = np.arange(65536, dtype=np.uint32)
    compass = [(m >> i) & 1 for i in (0, 4, 8, 12)]           # top, right, bottom, left
    any_c = (compass[0] | compass[1] | compass[2] | compass[3]).astype(bool)
    adj_c = ((compass[0] | compass[2]) & (compass[1] | compass[3])).astype(bool)
    three = (compass[0] + compass[1] + compass[2] + compass[3]) >= 3
    assert any_c[lut >= 4].all() and not any_c[lut >= 3].all()          # 4 is the smallest run length the ANY test may serve
    assert adj_c[lut >= 8].all() and not adj_c[lut >= 7].all()          # 8 the smallest for ADJ (a run of 7 can hold one compass point)
    assert three[lut >= 12].all() and not three[lut >= 11].all()        # the reference's kN >= 12 pre-check idea (fast.cpp:20-42)
    for diff in range(256):
        shift = 0
        while (2 << shift) <= diff + 1:
            shift += 1                                                   # 2^shift = largest power of two <= diff + 1 (fd_api.cu)
        mask = (0xFF << shift) & 0xFF
        a = np.arange(256)
        flagged = (a & mask) != 0
        assert flagged[a > diff].all(), diff                             # conservative: nothing above diff slips through
        if (diff + 1) & diff == 0:
            assert not flagged[a <= diff].any(), diff                    # exact for diff + 1 a power of two


def test_harris_pretest_threshold_is_the_exact_boundary():
    """The Harris kernel replaces the pre-test of feature_point_harris_detector.cpp:98, (trace * trace * 0.21f * inv_cnt2) > thr, by
    one compare against the smallest trace that passes (bisected on the host).  For many thresholds: that float passes the float32
    expression, the float just below it does not, and random traces agree with the expression."""
    lib = fd.load_library()
    inv = np.float32(1.0) / np.float32(9.0)
    inv2 = np.float32(inv * inv)

    def passes(trace, thr):
        t = np.float32(np.float32(np.float32(trace) * np.float32(trace)) * np.float32(0.21))
        return bool(np.float32(t * inv2) > np.float32(thr))

    rng = np.random.default_rng(3)
    thresholds = [0.0, 0.1, 1.0, 30.0, 1e3, 1e9, -1.0, 1e-30, 3.0e38] + rng.uniform(0, 500, 40).tolist() + (10 ** rng.uniform(-6, 12, 40)).tolist()
    out = C.c_float(0)
    with np.errstate(over="ignore"):
        for thr in thresholds:
            assert lib.fd_debug_harris_trace_min(C.c_float(thr), C.byref(out)) == 0
            t_min = np.float32(out.value)
            assert passes(t_min, thr), thr
            if t_min > 0:
                below = np.nextafter(t_min, np.float32(-1), dtype=np.float32)
                assert not passes(below, thr), thr
            for trace in (rng.uniform(0, 2, 50) * float(t_min) if np.isfinite(t_min) and t_min > 0 else rng.uniform(0, 1e6, 50)):
                assert passes(np.float32(trace), thr) == bool(np.float32(trace) >= t_min), (thr, trace)
        for thr in (float("inf"), float("nan")):        # nothing passes: the kernel's compare against NaN is false for every trace
            assert lib.fd_debug_harris_trace_min(C.c_float(thr), C.byref(out)) == 0 and np.isnan(out.value)
