import glob
import os
import sys
import warnings

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

warnings.filterwarnings("ignore", message=".*Sparse CSR tensor support.*")
warnings.filterwarnings("ignore", message=".*Sparse invariant checks.*")

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def golden_files(prefix):
    return sorted(glob.glob(os.path.join(GOLDEN, f"{prefix}_*.npz")))


def load_golden(path):
    z = np.load(path)
    return {k: z[k] for k in z.files}


def filter_dict(g):
    """Rebuild the {user_id: set(items)} dict stored flat in a LightGCN golden file."""
    out = {}
    uids = g["user_ids"]
    for r, i in zip(g["filter_rows"].tolist(), g["filter_items"].tolist()):
        out.setdefault(int(uids[r]), set()).add(int(i))
    return out


def assert_close(actual, expected, rtol=1e-5, atol_scale=1e-6, what=""):
    """allclose with atol tied to the tensor's magnitude (scores here are O(1e-4))."""
    actual = torch.as_tensor(actual).double().cpu()
    expected = torch.as_tensor(expected).double().cpu()
    assert actual.shape == expected.shape, f"{what}: shape {actual.shape} vs {expected.shape}"
    atol = atol_scale * float(expected.abs().max()) if expected.numel() else 0.0
    err = (actual - expected).abs()
    bound = atol + rtol * expected.abs()
    bad = err > bound
    assert not bool(bad.any()), (
        f"{what}: {int(bad.sum())} of {bad.numel()} outside rtol={rtol} atol={atol:.3e}; "
        f"max err {float(err.max()):.3e}")
