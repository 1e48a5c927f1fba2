"""pytest configuration: the `gpu` marker, import paths, shared helpers."""
import importlib.util
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "juicy-audio-plugins_b200")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def load_juicy_batch():
    """The package directory name carries hyphens, so load its module by path."""
    if "juicy_batch" in sys.modules:
        return sys.modules["juicy_batch"]
    spec = importlib.util.spec_from_file_location("juicy_batch", os.path.join(PKG, "juicy_batch.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["juicy_batch"] = mod
    spec.loader.exec_module(mod)
    return mod


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    # gpu tests fail loudly (not skip) on a box whose GPU or extension is missing when selected
    # with -m gpu; in a CPU-only run without -m they are deselected by the driver's -m "not gpu".
    pass


@pytest.fixture(scope="session")
def jb():
    return load_juicy_batch()


@pytest.fixture(scope="session")
def port():
    from oracle import port as p
    p.lib()
    return p


@pytest.fixture(scope="session")
def refhost():
    from oracle import refhost as r
    return r


SAMPLE_TOL = 1.0e-5   # |gpu - ref| <= SAMPLE_TOL * max|ref| over the clip (BASELINE.json north_star)
METRIC_TOL = 0.01     # absolute, on the 0..100 score / juiciness and on every 0..1 feature bar
# record fields on a 0..100 scale share the same absolute tolerance as the north_star states it for the score
FIELDS_0_100 = (0, 1, 2, 13, 14)


def assert_samples_close(got, ref, what=""):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    assert np.isfinite(got).all(), "%s: non-finite output" % what
    peak = np.abs(ref).max(axis=tuple(range(1, ref.ndim)), keepdims=True) if ref.ndim > 1 else np.abs(ref).max()
    peak = np.maximum(peak, 1.0e-30)
    err = np.abs(got - ref) / peak
    worst = float(err.max()) if err.size else 0.0
    assert worst <= SAMPLE_TOL, "%s: max |gpu-ref|/peak = %.3e > %.1e" % (what, worst, SAMPLE_TOL)
    return worst


def assert_records_close(got, ref, what=""):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    err = np.abs(got - ref)
    worst = float(err.max()) if err.size else 0.0
    if worst > METRIC_TOL:
        idx = np.unravel_index(int(err.argmax()), err.shape)
        raise AssertionError("%s: metric record differs by %.4f at %s (got %.5f, ref %.5f)"
                             % (what, worst, idx, got[idx], ref[idx]))
    return worst
