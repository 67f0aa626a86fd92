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
        # a kernel that never returns must end the run, not hang the box: hard per-test limit for GPU tests
        # (the "thread" method works even while the interpreter is blocked inside a CUDA call)
        for item in items:
            if "gpu" in item.keywords and item.get_closest_marker("timeout") is None:
                item.add_marker(pytest.mark.timeout(300, method="thread"))
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


@pytest.fixture(scope="session")
def hnm_lib():
    """The built C-ABI library (compiled in-tree on first use; no GPU needed to build or load)."""
    from hnm_recommendation_b200 import _lib, build
    build.build_library()
    return _lib.load()


def assert_topk_matches_scores(got_ids, ref_scores, k, rel_tol=1e-6):
    """Near-tie-aware top-k check against a full fp64 reference score matrix.

    Row r passes when got_ids[r] are distinct, and ranking them by the reference scores
    is as good as the true ranking position by position up to rel_tol * max|score| --
    i.e. a mismatch against the canonical list can only involve items whose reference
    scores differ by less than the tolerance.  Returns the number of rows that differ
    from the canonical (score desc, id asc) list at all.
    """
    got = torch.as_tensor(got_ids).cpu()
    s = torch.as_tensor(ref_scores).double().cpu()
    canon = torch.sort(s, dim=1, descending=True, stable=True)
    want_ids, want_s = canon.indices[:, :k], canon.values[:, :k]
    assert got.shape == want_ids.shape
    differ = (got != want_ids).any(dim=1)
    if differ.any():
        rows = differ.nonzero().view(-1)
        g = torch.gather(s[rows], 1, got[rows])
        finite = torch.where(torch.isfinite(s[rows]), s[rows].abs(), torch.zeros_like(s[rows]))
        tol = rel_tol * finite.max(dim=1, keepdim=True).values
        w = want_s[rows]
        ok = (g >= w - tol) | (torch.isinf(w) & torch.isinf(g))
        assert bool(ok.all()), f"{int((~ok).any(dim=1).sum())} rows differ beyond near-ties"
        for r in range(rows.numel()):
            assert len(set(got[rows[r]].tolist())) == k, "duplicate ids in a top-k row"
    return int(differ.sum())
