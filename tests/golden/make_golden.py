"""Generate golden vectors by running the REFERENCE's own model code.

Run in the authoring container only (needs /root/reference; never at test time):

    python tests/golden/make_golden.py

The reference cannot be imported as shipped: three
third-party packages are not installed and one class is defined nowhere.  This
script supplies the minimum stand-ins and then executes the unmodified files
``/root/reference/src/models/lightgcn.py`` and ``neural_cf.py``:

  pytorch_lightning.LightningModule  -> torch.nn.Module + no-op save_hyperparameters()/log()
  src.evaluation.RecommendationMetrics -> empty class (never touched on the scoring path)
  torch_sparse.sum(src, index, dim, dim_size) -> scatter-add of src by index
        (the call at lightgcn.py:103 uses torch_scatter's signature; the intended
        meaning deg[i] = sum of w over edges with row == i is unambiguous)
  torch_sparse.SparseTensor(row, col, value, sparse_sizes) with ``@ dense``
        -> storage sorted by (row, col), duplicates retained, matmul = sum over a
        row's entries of value * dense[col]  (rusty1s/pytorch_sparse public semantics)

Everything else -- self loops, degree normalisation, alpha weights, layer sum
order, gather, matmul, filtering -- is the reference's code running unchanged.
``torch.topk`` ties are implementation-defined, so golden top-k lists are stored
through a stable descending sort of the reference's score matrix (BASELINE.json
tie-break) and the raw ``recommend()`` output is stored beside them for cases
without ties.
"""
import importlib
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def install_stubs():
    pl = types.ModuleType("pytorch_lightning")

    class LightningModule(torch.nn.Module):
        def save_hyperparameters(self, *a, **k):
            pass

        def log(self, *a, **k):
            pass

    pl.LightningModule = LightningModule
    sys.modules["pytorch_lightning"] = pl

    ts = types.ModuleType("torch_sparse")

    def ts_sum(src, index, dim=0, dim_size=None):
        out = torch.zeros(dim_size, dtype=src.dtype, device=src.device)
        return out.scatter_add_(dim, index, src)

    class SparseTensor:
        def __init__(self, row, col, value, sparse_sizes):
            n = sparse_sizes[1]
            order = torch.argsort(row * n + col, stable=True)
            self.row, self.col, self.value = row[order], col[order], value[order]
            self.sizes = sparse_sizes

        def __matmul__(self, dense):
            out = torch.zeros(self.sizes[0], dense.size(1), dtype=dense.dtype)
            return out.index_add_(0, self.row, self.value.unsqueeze(1) * dense[self.col])

    ts.sum = ts_sum
    ts.SparseTensor = SparseTensor
    sys.modules["torch_sparse"] = ts

    # package skeleton so that `from ..evaluation import RecommendationMetrics` resolves
    src = types.ModuleType("src"); src.__path__ = [os.path.join(REF, "src")]
    models = types.ModuleType("src.models"); models.__path__ = [os.path.join(REF, "src", "models")]
    ev = types.ModuleType("src.evaluation")

    class RecommendationMetrics:
        def __init__(self, top_k=12):
            self.top_k = top_k

    ev.RecommendationMetrics = RecommendationMetrics
    sys.modules.update({"src": src, "src.models": models, "src.evaluation": ev})


def load_reference():
    install_stubs()
    lg = importlib.import_module("src.models.lightgcn")
    ncf = importlib.import_module("src.models.neural_cf")
    return lg.LightGCN, ncf.NeuralCF


def random_bipartite(num_users, num_items, num_edges, gen):
    """Edge list in the layout of the reference's own test (tests/test_models.py:178-185)."""
    u = torch.randint(0, num_users, (num_edges,), generator=gen)
    i = torch.randint(0, num_items, (num_edges,), generator=gen) + num_users
    return torch.stack([torch.cat([u, i]), torch.cat([i, u])])


LIGHTGCN_CASES = [
    # name, U, I, d, L, k, E, alpha, weighted, seed
    ("ref_fixture", 100, 50, 16, 3, 5, 200, None, False, 1),     # tests/test_models.py:14-22,178-187
    ("alpha_decay", 100, 50, 16, 3, 5, 200, 0.5, False, 2),
    ("weighted", 80, 40, 32, 2, 12, 300, None, True, 3),
    ("deep_dupes", 60, 30, 64, 4, 12, 2000, None, False, 4),      # many duplicate edges
    ("sparse_isolated", 300, 200, 64, 3, 12, 150, None, False, 5),  # most nodes isolated
    ("one_layer", 50, 70, 8, 1, 12, 400, 0.3, True, 6),
    ("mid", 400, 300, 64, 3, 12, 6000, None, False, 7),
]


def make_lightgcn(LightGCN):
    for name, U, I, d, L, k, E, alpha, weighted, seed in LIGHTGCN_CASES:
        gen = torch.Generator().manual_seed(seed)
        torch.manual_seed(seed)
        model = LightGCN(num_users=U, num_items=I, embedding_dim=d, num_layers=L, top_k=k, alpha=alpha)
        edge_index = random_bipartite(U, I, E, gen)
        ew = None
        if weighted:
            half = torch.rand(E, generator=gen) * 2 + 0.25
            ew = torch.cat([half, half])
        model.set_graph(edge_index, ew)
        with torch.no_grad():
            ue, ie = model.forward()
            uids = torch.randperm(U, generator=gen)[:min(U, 48)]
            scores = model.predict_all_items(uids)
            iids = torch.randint(0, I, (uids.numel(),), generator=gen)
            pair = model.predict(uids, iids)
            rec = model.recommend(uids)
            filt = {int(u): set(torch.randint(0, I, (7,), generator=gen).tolist())
                    for u in uids[::3].tolist()}
            rec_f = model.recommend(uids, filter_items=filt)
            s_f = scores.clone()
            for r, u in enumerate(uids.tolist()):
                if u in filt:
                    s_f[r, list(filt[u])] = float("-inf")
        canon = torch.sort(scores, dim=1, descending=True, stable=True).indices[:, :k]
        canon_f = torch.sort(s_f, dim=1, descending=True, stable=True).indices[:, :k]
        filt_rows = np.array([r for r, u in enumerate(uids.tolist()) if u in filt for _ in filt[u]], dtype=np.int64)
        filt_items = np.array([i for u in uids.tolist() if u in filt for i in sorted(filt[u])], dtype=np.int64)
        np.savez_compressed(
            os.path.join(OUT, f"lightgcn_{name}.npz"),
            num_users=U, num_items=I, embedding_dim=d, num_layers=L, top_k=k,
            alpha=np.nan if alpha is None else alpha,
            alpha_list=np.array(model.alpha, dtype=np.float64),
            edge_index=edge_index.numpy(),
            edge_weight=np.zeros(0, np.float32) if ew is None else ew.numpy(),
            weight=model.embeddings.weight.detach().numpy(),
            user_emb=ue.numpy(), item_emb=ie.numpy(),
            user_ids=uids.numpy(), item_ids=iids.numpy(),
            scores=scores.numpy(), pair_scores=pair.numpy(),
            recommend_raw=rec.numpy(), recommend_filtered_raw=rec_f.numpy(),
            topk_canonical=canon.numpy(), topk_filtered_canonical=canon_f.numpy(),
            filter_rows=filt_rows, filter_items=filt_items,
        )
        print("lightgcn", name, "ok", tuple(ue.shape), tuple(ie.shape))


NCF_CASES = [
    # name, U, I, mf_dim, mlp_dims, k, seed
    ("default", 120, 90, 64, [128, 64, 32], 12, 11),   # configs/model/neural_cf.yaml:5-15
    ("small", 40, 1100, 8, [16, 8], 5, 12),            # >1000 items: exercises the chunk loop :167
    ("deep", 64, 48, 16, [64, 32, 16, 8], 12, 13),
]


def make_ncf(NeuralCF):
    for name, U, I, mf, dims, k, seed in NCF_CASES:
        gen = torch.Generator().manual_seed(seed)
        torch.manual_seed(seed)
        model = NeuralCF(num_users=U, num_items=I, mf_dim=mf, mlp_dims=dims, top_k=k)
        # non-zero biases so the bias path is exercised (the reference zero-inits them)
        with torch.no_grad():
            for p_name, p in model.named_parameters():
                if p_name.endswith("bias"):
                    p.copy_(torch.randn(p.shape, generator=gen) * 0.05)
        model.eval()
        with torch.no_grad():
            uids = torch.randint(0, U, (200,), generator=gen)
            iids = torch.randint(0, I, (200,), generator=gen)
            logits = model(uids, iids)
            single = model(uids[:1], iids[:1])
            au = torch.randperm(U, generator=gen)[:16]
            all_scores = model.predict_all_items(au)
            rec = model.recommend(au)
        canon = torch.sort(all_scores, dim=1, descending=True, stable=True).indices[:, :k]
        state = {f"state.{n}": v.detach().numpy() for n, v in model.state_dict().items()}
        np.savez_compressed(
            os.path.join(OUT, f"ncf_{name}.npz"),
            num_users=U, num_items=I, mf_dim=mf, mlp_dims=np.array(dims), top_k=k,
            user_ids=uids.numpy(), item_ids=iids.numpy(), logits=logits.numpy(),
            single_logit=single.numpy(), all_user_ids=au.numpy(), all_scores=all_scores.numpy(),
            recommend_raw=rec.numpy(), topk_canonical=canon.numpy(), **state)
        print("ncf", name, "ok", tuple(logits.shape), "single dim", single.dim())


if __name__ == "__main__":
    LightGCN, NeuralCF = load_reference()
    make_lightgcn(LightGCN)
    make_ncf(NeuralCF)
