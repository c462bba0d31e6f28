"""Host logic of the drop-in line detector on the CPU: feature_detector_b200/cpp/line_segments_host.cpp (region growing,
rectangle fit, validation -- the stage the north star leaves on the host) fed by the oracle's dense stage instead of kernel 5
(tests/hoststage/hoststage_check.cpp replaces the one GPU-calling member function), against the reference's own segments:
committed golden vectors (tests/golden/lsd_segments.npz, made by make_golden_lsd.py from oracle/_ref) and, where the reference
build is present, the reference run live."""
import importlib.util
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
BUILD = os.path.join(ROOT, "tests", "_build")


def _cases():
    spec = importlib.util.spec_from_file_location("make_golden_lsd", os.path.join(GOLDEN, "make_golden_lsd.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="module")
def hoststage(built):
    os.makedirs(BUILD, exist_ok=True)
    exe = os.path.join(BUILD, "hoststage_check")
    cpp = os.path.join(ROOT, "feature_detector_b200", "cpp")
    cmd = ["g++", "-std=c++17", "-O2", "-g", "-Wall", "-I" + cpp, "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "compat", "slam_utility"),
           "-o", exe, os.path.join(ROOT, "tests", "hoststage", "hoststage_check.cpp"), os.path.join(cpp, "line_segments_host.cpp"),
           "-L" + os.path.join(ROOT, "oracle"), "-lfd_oracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle")]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-4000:]

    def run(frame, needed, min_norm, host_libm=True, seeds=None, tmp=os.path.join(BUILD, "hoststage_frame.u8")):
        frame = np.ascontiguousarray(frame, np.uint8)
        frame.tofile(tmp)
        cmd = [exe, tmp, str(frame.shape[0]), str(frame.shape[1]), str(needed), repr(float(min_norm)), "1" if host_libm else "0"]
        if seeds is not None:   # (n, 2) row / col pairs replacing the oracle's seed order
            np.ascontiguousarray(seeds, np.int32).tofile(tmp + ".seeds")
            cmd.append(tmp + ".seeds")
        r = subprocess.run(cmd, stdout=subprocess.PIPE, text=True)
        head, *rows = r.stdout.strip().splitlines()
        info = dict(zip(head.split()[::2], (int(v) for v in head.split()[1::2])))
        lines = np.array([[int(x, 16) for x in row.split()] for row in rows], np.uint32).reshape(-1, 4).view(np.float32)
        return r.returncode, info, lines
    return run


@pytest.mark.parametrize("name", ["image", "synth752", "synth_odd", "synth_norm35"])
def test_host_stage_returns_the_reference_segments(hoststage, port, name):
    """Fed the reference's seed order, the host stage returns the reference's segments: same number, order and bits.  (The seed order
    is an input here because the reference's std::sort is unstable: equal-norm seeds come out in the C++ library's order, and which of
    them grows first decides the segments.  The oracle port and the GPU kernel keep ties in push order; the test below pins that this
    is the only difference.)"""
    mod = _cases()
    gold = np.load(os.path.join(GOLDEN, "lsd_segments.npz"))
    _, needed, min_norm = mod.CASES[name]
    frame = mod.frame_of(name)
    ref_seeds = gold[name + ".seeds"].astype(np.int32)
    for host_libm in (True, False):   # angles recomputed with the host libm, or taken from the dense stage as they come
        rc, info, lines = hoststage(frame, needed, min_norm, host_libm, ref_seeds)
        assert rc == 0 and info["ok"] == 1 and info["lines"] == info["rectangles"] == len(gold[name]) > 0 and info["seeds"] == len(ref_seeds)
        assert np.array_equal(lines.view(np.uint32), gold[name].view(np.uint32))
    # the port's seed order is the reference's up to the order inside runs of equal norm
    m = port.lsd_map(frame, min_norm)
    mine = m["sorted_rc"]
    assert len(mine) == len(ref_seeds)
    n_mine, n_ref = m["norm"][mine[:, 0], mine[:, 1]], m["norm"][ref_seeds[:, 0], ref_seeds[:, 1]]
    assert np.array_equal(n_mine, n_ref) and np.all(np.diff(n_mine) <= 0)
    key = lambda rc: rc[:, 1].astype(np.int64) * frame.shape[0] + rc[:, 0]
    assert np.array_equal(np.sort(key(mine)), np.sort(key(ref_seeds)))                       # same set of seeds
    assert np.all(np.diff(key(mine))[np.diff(n_mine) == 0] > 0)                              # ties: column outer, row inner (push order)
    # with its own (stable) order the host stage still finds segments; identical to the reference's whenever no tie matters
    rc, info, own = hoststage(frame, needed, min_norm)
    assert rc == 0 and info["lines"] > 0
    if name in ("image", "synth_odd"):
        assert np.array_equal(own.view(np.uint32), gold[name].view(np.uint32))
    from oracle.bindings import Ref, have_ref
    if have_ref():   # and live, where the reference build is present
        ok, ref_lines = Ref().lsd_detect(frame, needed, min_norm)
        assert ok and np.array_equal(gold[name].view(np.uint32), ref_lines.view(np.uint32))
        assert np.array_equal(Ref().lsd_map(frame, min_norm)["sorted_rc"], ref_seeds)


def test_host_stage_edge_cases(hoststage):
    flat = np.full((60, 80), 128, np.uint8)
    rc, info, lines = hoststage(flat, 200, 20.0)
    assert rc == 0 and info["ok"] == 1 and info["lines"] == 0 and info["seeds"] == 0           # nothing valid: true, no segments
    rc, info, lines = hoststage(_cases().frame_of("synth_odd"), 0, 20.0)
    assert rc == 0 and info["ok"] == 1 and info["lines"] == 0                                  # needed == 0 returns early (.cpp:15)
    rc, info, lines = hoststage(np.zeros((1, 50), np.uint8), 10, 20.0)
    assert rc == 1 and info["ok"] == 0                                                         # fewer than two rows: false (.cpp:14)
    # a perfectly vertical step gives a region with Ixy == 0, which the reference's rectangle fit abandons (.cpp:186-188): no segment
    step = np.zeros((120, 160), np.uint8)
    step[:, 80:] = 200
    rc, info, lines = hoststage(step, 10, 20.0)
    assert rc == 0 and info["ok"] == 1 and info["seeds"] > 0 and info["lines"] == 0
    # a soft edge with a little texture: one long segment along it (the reference finds (94.0, 1.5) - (94.0, 157.5))
    edge = (np.clip((np.arange(200)[None, :] - 90) * 30, 0, 255) + (np.arange(160)[:, None] % 7)).astype(np.uint8)
    rc, info, lines = hoststage(edge, 10, 20.0)
    assert rc == 0 and info["lines"] == 1
    x0, y0, x1, y1 = lines[0]
    assert abs(x0 - 94) < 1 and abs(x1 - 94) < 1 and abs(abs(y1 - y0) - 156) < 3
