"""configs[4] (5 M users x 1 M items, d = 256, 4 layers) on one GPU: how the nomination margins of the fused
scorer trade candidate-list size against the share of users the certificate rejects.

    python tools/sweep_margin.py [scale]            # scale < 1 shrinks users/items/interactions together
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hnm_recommendation_b200 import synth                      # noqa: E402
from hnm_recommendation_b200.lightgcn import LightGCN          # noqa: E402
from hnm_recommendation_b200.scorer import FusedScorer         # noqa: E402


def main():
    f = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
    u, i, e = int(5_000_000 * f), int(1_000_000 * f), int(200_000_000 * f)
    dev = torch.device("cuda:0")
    model = LightGCN(u, i, embedding_dim=256, num_layers=4, top_k=12).to(dev)
    with torch.no_grad():
        model.embeddings.weight.normal_(0.0, 0.1, generator=torch.Generator(device=dev).manual_seed(42))
    model.set_graph(synth.interactions(u, i, e, seed=42).edge_index().to(dev))
    model.eval()
    with torch.no_grad():
        ue, ie = model.forward()
    for m1, m2, cap in ((3, 12, 384), (4, 12, 384), (5, 12, 384), (6, 12, 384), (5, 12, 320)):
        sc = FusedScorer(ue, ie, sel_margin=m1, tier2_margin=m2, cand_cap=cap)
        sc.topk(None, 12)
        a = torch.cuda.Event(enable_timing=True)
        b = torch.cuda.Event(enable_timing=True)
        a.record()
        sc.topk(None, 12)
        b.record()
        torch.cuda.synchronize()
        whole = a.elapsed_time(b)
        stats_fast = dict(sc.last_stats)
        sc.profile = True
        sc.topk(None, 12)
        print(json.dumps({"margin": m1, "tier2_margin": m2, "cand_cap": cap, "topk_ms": whole, "stats": stats_fast,
                          "stages_ms": sc.stage_ms, "why": sc.last_stats}), flush=True)
        del sc
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
