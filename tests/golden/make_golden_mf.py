"""Golden vectors for MatrixFactorization (src/models/matrix_factorization.py), produced by running the
REFERENCE's own, unmodified file with the stand-ins of make_golden.py (pytorch_lightning.LightningModule ->
nn.Module, the undefined RecommendationMetrics -> empty class).  Authoring container only:

    python tests/golden/make_golden_mf.py
"""
import importlib
import os

import numpy as np
import torch

import make_golden as G

OUT = os.path.dirname(os.path.abspath(__file__))
CASES = [  # name, U, I, d, k, seed
    ("default", 150, 700, 64, 12, 31),        # configs/model/matrix_factorization.yaml: embedding_dim 64
    ("small_dim", 90, 400, 16, 5, 32),
]


def main():
    G.install_stubs()
    MF = importlib.import_module("src.models.matrix_factorization").MatrixFactorization
    for name, U, I, d, k, seed in CASES:
        gen = torch.Generator().manual_seed(seed)
        torch.manual_seed(seed)
        model = MF(num_users=U, num_items=I, embedding_dim=d, top_k=k, sparse=False)
        with torch.no_grad():                  # the reference zero-inits the biases: give them values so they matter
            model.user_bias.weight.copy_(torch.randn(U, 1, generator=gen) * 0.01)
            model.item_bias.weight.copy_(torch.randn(I, 1, generator=gen) * 0.01)
            model.global_bias.copy_(torch.tensor([0.003]))
        model.eval()
        with torch.no_grad():
            uids = torch.randint(0, U, (120,), generator=gen)
            iids = torch.randint(0, I, (120,), generator=gen)
            pred = model(uids, iids)
            au = torch.randperm(U, generator=gen)[:32]
            scores = model.predict_all_items(au)
            rec = model.recommend(au)
            filt = {int(u): set(torch.randint(0, I, (9,), generator=gen).tolist()) for u in au[::2].tolist()}
            rec_f = model.recommend(au, filter_items=filt)
            s_f = scores.clone()
            for r, u in enumerate(au.tolist()):
                if u in filt:
                    s_f[r, list(filt[u])] = float("-inf")
        canon = torch.sort(scores, dim=1, descending=True, stable=True).indices[:, :k]
        canon_f = torch.sort(s_f, dim=1, descending=True, stable=True).indices[:, :k]
        filt_rows = np.array([r for r, u in enumerate(au.tolist()) if u in filt for _ in filt[u]], dtype=np.int64)
        filt_items = np.array([i for u in au.tolist() if u in filt for i in sorted(filt[u])], dtype=np.int64)
        state = {f"state.{n}": v.detach().numpy() for n, v in model.state_dict().items()}
        np.savez_compressed(os.path.join(OUT, f"mf_{name}.npz"), num_users=U, num_items=I, embedding_dim=d, top_k=k,
                            user_ids=uids.numpy(), item_ids=iids.numpy(), pred=pred.numpy(), all_user_ids=au.numpy(),
                            scores=scores.numpy(), recommend_raw=rec.numpy(), recommend_filtered_raw=rec_f.numpy(),
                            topk_canonical=canon.numpy(), topk_filtered_canonical=canon_f.numpy(),
                            filter_rows=filt_rows, filter_items=filt_items, **state)
        print("mf", name, "ok", tuple(scores.shape))


if __name__ == "__main__":
    main()
