import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def rt():
    """The product package (hyphenated directory name -> importlib)."""
    return importlib.import_module("html5-canvas-raytracer_b200")


@pytest.fixture(scope="session")
def oracle():
    """CPU float64 restatement of main.js — the checker, never the thing under test."""
    from oracle import oracle as o
    o.build()
    return o


def lsb_stats(a: np.ndarray, b: np.ndarray):
    """Per-PIXEL comparison of two RGBA8 frames: fraction of pixels whose every channel is within
    1 / 2 LSB, max abs diff, PSNR over RGB (dB)."""
    a = np.asarray(a).astype(np.int32)
    b = np.asarray(b).astype(np.int32)
    assert a.shape == b.shape
    d = np.abs(a - b)[..., :3].reshape(-1, 3)
    worst = d.max(axis=1)
    mse = float((d.astype(np.float64) ** 2).mean())
    return {
        "exact": float((worst == 0).mean()),
        "within1": float((worst <= 1).mean()),
        "within2": float((worst <= 2).mean()),
        "max": int(worst.max()) if worst.size else 0,
        "psnr": float("inf") if mse == 0 else 10.0 * np.log10(255.0 ** 2 / mse),
        "alpha_ok": bool((a[..., 3] == b[..., 3]).all()),
    }


# Parity gates of BASELINE.json's north_star.
FP64_WITHIN1 = 0.9999   # fp64 validation kernel: +-1 LSB per channel on >= 99.99 % of pixels
FP32_WITHIN2 = 0.999    # fp32 throughput kernel: +-2 LSB on >= 99.9 % of pixels
FP64_PSNR_DB = 50.0     # stated PSNR floors (a checker flip is a 255-step error)
FP32_PSNR_DB = 40.0
