"""Host logic behind the line detector's dense stage on the CPU: tests/hoststage/line_segments_host.cpp (test scaffolding: region growing,
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
    cmd = ["g++", "-std=c++17", "-O2", "-g", "-Wall", "-I" + cpp, "-I" + os.path.join(ROOT, "tests", "hoststage"), "-I" + os.path.join(ROOT, "include"),
           "-I" + os.path.join(ROOT, "compat", "slam_utility"),
           "-o", exe, os.path.join(ROOT, "tests", "hoststage", "hoststage_check.cpp"), os.path.join(ROOT, "tests", "hoststage", "line_segments_host.cpp"),
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


def test_host_glue_is_clean_under_asan_and_ubsan(built, image_png, tmp_path):
    """SURVEY.md section 5 (the reference builds without sanitizers): the drop-in classes' host code under -fsanitize=address,undefined,
    with the oracle behind the C ABI -- the whole fd_dropin_check replay, and the line detector on frames down to 2 x 2 pixels (a
    two-column frame makes the reference itself write out of bounds in feature_line_detector.cpp:64-68; the drop-in guards that line)."""
    cpp = os.path.join(ROOT, "feature_detector_b200", "cpp")
    os.makedirs(BUILD, exist_ok=True)
    flags = ["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-I" + cpp, "-I" + os.path.join(ROOT, "tests", "hoststage"), "-I" + os.path.join(ROOT, "include"),
             "-I" + os.path.join(ROOT, "compat", "slam_utility")]
    link = ["-L" + os.path.join(ROOT, "oracle"), "-lfd_oracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle")]
    hs = os.path.join(BUILD, "hoststage_check_asan")
    full = os.path.join(BUILD, "fd_dropin_check_asan")
    probe = subprocess.run(flags + ["-x", "c++", "-", "-o", os.path.join(BUILD, "asan_probe")], input="int main(){return 0;}", text=True,
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    if probe.returncode != 0:
        pytest.skip("this toolchain has no sanitizer runtime")
    hsdir = os.path.join(ROOT, "tests", "hoststage")
    for exe, srcs in ((hs, [os.path.join(hsdir, "hoststage_check.cpp"), os.path.join(hsdir, "line_segments_host.cpp")]),
                      (full, [os.path.join(cpp, f) for f in ("fd_dropin_check.cpp", "feature_point_detector.cpp", "descriptor_brief.cpp", "feature_line_field.cpp",
                                                             "nn_feature_point_postprocess.cpp")] +
                             [os.path.join(hsdir, "line_segments_host.cpp"), os.path.join(hsdir, "fake_fd_abi.cpp")])):
        r = subprocess.run(flags + ["-o", exe] + srcs + link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        assert r.returncode == 0, r.stdout[-3000:]
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=1", UBSAN_OPTIONS="halt_on_error=1:print_stacktrace=1")

    def clean(cmd):
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env, timeout=600)
        assert r.returncode == 0 and "ERROR" not in r.stderr and "runtime error" not in r.stderr, (cmd, r.stderr[-3000:])
        return r.stdout

    rng = np.random.default_rng(0)
    from feature_detector_b200.synth import synth, synth_descriptor_volume, synth_heatmap
    for shape in [(2, 2), (2, 5), (5, 2), (3, 3), (4, 4), (6, 2), (2, 9), (7, 7), (60, 80), (217, 333)]:
        frame = rng.integers(0, 256, shape, dtype=np.uint8) if shape[0] < 50 else synth(shape[1], shape[0], 3)
        path = tmp_path / "frame.u8"
        frame.tofile(path)
        out = clean([hs, str(path), str(shape[0]), str(shape[1]), "50", "5.0"])
        assert out.startswith("ok 1")
    for shape in [(1, 1), (2, 2), (3, 9), (7, 7), (10, 200), (41, 41), (64, 40)]:      # the whole replay (detectors, BRIEF, line detector) on small frames
        path = tmp_path / "small.u8"
        rng.integers(0, 256, shape, dtype=np.uint8).tofile(path)
        clean([full, str(path), str(shape[0]), str(shape[1])])
    raw = tmp_path / "image.u8"
    raw.write_bytes(image_png.tobytes())
    synth_heatmap(752, 480, 7).tofile(tmp_path / "heat.f32")
    synth_descriptor_volume(256, 60, 94, 7).tofile(tmp_path / "vol.f32")
    out = clean([full, str(raw), "480", "752", str(tmp_path / "heat.f32"), str(tmp_path / "vol.f32"), "256", "5"])
    assert '"lsd_detect": {"ok": true, "n_lines": 40' in out
