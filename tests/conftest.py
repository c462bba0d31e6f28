import gzip
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

SURVEY_FNV_BASIS = 1469598103934665603  # the (truncated) basis SURVEY.md section 8c hashed with


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def fnv(arr, basis=SURVEY_FNV_BASIS):
    h = basis
    for b in np.ascontiguousarray(arr).tobytes():
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return "%016x" % h


def _run(cmd, cwd):
    r = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"{' '.join(cmd)} failed:\n{r.stdout[-4000:]}")


@pytest.fixture(scope="session", autouse=True)
def built():
    """Make sure the native pieces exist (no-ops when __graft_entry__.build() already ran)."""
    if not os.path.exists(os.path.join(ROOT, "feature_detector_b200", "libfd_b200.so")):
        _run(["make", "-j8", "-C", os.path.join(ROOT, "feature_detector_b200", "csrc")], ROOT)
    if not os.path.exists(os.path.join(ROOT, "oracle", "libfd_oracle.so")):
        _run(["make", "-C", os.path.join(ROOT, "oracle"), "libfd_oracle.so"], ROOT)
    return True


@pytest.fixture(scope="session")
def image_png():
    """Raw decode of the reference's examples/image.png (752x480 u8), committed as a fixture."""
    with gzip.open(os.path.join(GOLDEN, "image_752x480.u8.gz"), "rb") as f:
        return np.frombuffer(f.read(), np.uint8).reshape(480, 752).copy()


@pytest.fixture(scope="session")
def kat():
    with open(os.path.join(GOLDEN, "kat.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def vectors():
    return dict(np.load(os.path.join(GOLDEN, "vectors.npz")))


@pytest.fixture(scope="session")
def frames(image_png):
    from feature_detector_b200.synth import synth
    return {"image": image_png, "synth752": synth(752, 480, 0), "synth_odd": synth(333, 217, 5)}


@pytest.fixture(scope="session")
def port(built):
    from oracle.bindings import Port
    return Port()


@pytest.fixture(scope="session")
def ref():
    from oracle.bindings import Ref, have_ref
    if not have_ref():
        pytest.skip("oracle/_ref/libfd_ref.so not built here (needs /root/reference)")
    return Ref()


@pytest.fixture(scope="session")
def checker(built):
    """The strongest CPU checker available: the reference compiled in place, else the C port."""
    from oracle.bindings import best_checker
    return best_checker()
