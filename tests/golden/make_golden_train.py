"""Golden vectors for the training-side entry point: the REFERENCE's own ``LightGCN.bpr_loss``
(src/models/lightgcn.py:206-245) and the gradient autograd derives from it.

Run in the authoring container only (needs /root/reference; never at test time):

    python tests/golden/make_golden_train.py

Uses the same stand-ins as make_golden.py (see its docstring); the reference file runs unchanged.
"""
import os

import numpy as np
import torch

from make_golden import OUT, load_reference, random_bipartite

CASES = [
    # name, U, I, d, L, E, alpha, weighted, weight_decay, batch, seed
    ("bpr_default", 90, 40, 64, 3, 700, None, False, 1e-4, 128, 21),
    ("bpr_weighted_alpha", 60, 50, 32, 2, 500, 0.5, True, 1e-2, 64, 22),
]


def main():
    LightGCN, _ = load_reference()
    for name, U, I, d, L, E, alpha, weighted, wd, B, seed in CASES:
        gen = torch.Generator().manual_seed(seed)
        torch.manual_seed(seed)
        model = LightGCN(num_users=U, num_items=I, embedding_dim=d, num_layers=L, alpha=alpha, weight_decay=wd)
        with torch.no_grad():
            model.embeddings.weight.mul_(8.0)          # Xavier on a tiny table is too flat for a useful gradient check
        edge_index = random_bipartite(U, I, E, gen)
        ew = None
        if weighted:
            half = torch.rand(E, generator=gen) * 2 + 0.25
            ew = torch.cat([half, half])
        model.set_graph(edge_index, ew)
        uid = torch.randint(0, U, (B,), generator=gen)
        pos = torch.randint(0, I, (B,), generator=gen)
        neg = torch.randint(0, I, (B,), generator=gen)
        loss = model.training_step({"user_ids": uid, "pos_items": pos, "neg_items": neg}, 0)
        loss.backward()
        np.savez_compressed(
            os.path.join(OUT, f"train_{name}.npz"),
            num_users=U, num_items=I, embedding_dim=d, num_layers=L, weight_decay=wd,
            alpha=np.nan if alpha is None else alpha,
            edge_index=edge_index.numpy(), edge_weight=np.zeros(0, np.float32) if ew is None else ew.numpy(),
            weight=model.embeddings.weight.detach().numpy(),
            user_ids=uid.numpy(), pos_items=pos.numpy(), neg_items=neg.numpy(),
            loss=np.float64(loss.item()), grad=model.embeddings.weight.grad.numpy())
        print("train", name, "loss", float(loss), "grad absmax", float(model.embeddings.weight.grad.abs().max()))


if __name__ == "__main__":
    main()
