"""pytest configuration: the `gpu` marker and shared fixture loaders."""
import json
import os
import struct
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _ensure_library():
    """The CUDA library is a build artefact (git-ignored): build it once if this checkout does not have it yet."""
    so = os.path.join(ROOT, "cmu-11785-idl-1.58bit-asr_b200", "libonebit.so")
    if not os.path.exists(so):
        import subprocess
        subprocess.run(["bash", os.path.join(ROOT, "cmu-11785-idl-1.58bit-asr_b200", "csrc", "build.sh")], check=True)


_ensure_library()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(autouse=True)
def _restore_backend_flags():
    """Tests that pin torch's own kernels to true fp32 must not leak that choice into later tests."""
    import torch
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.deterministic)
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.deterministic = saved


def bits_to_f32(s: str) -> np.float32:
    return np.float32(struct.unpack("<f", struct.pack("<I", int(s, 16)))[0])


@pytest.fixture(scope="session")
def kat_edges():
    with open(os.path.join(GOLDEN, "kat_edges.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def kat_seeded():
    with open(os.path.join(GOLDEN, "kat_seeded.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def kat_layer():
    return dict(np.load(os.path.join(GOLDEN, "kat_layer.npz")))


@pytest.fixture(scope="session")
def kat_layer_stats():
    with open(os.path.join(GOLDEN, "kat_layer_stats.json")) as f:
        return json.load(f)
