"""Dev probe: fused kernel alone at H&M shape under HNM_FUSED_DEBUG modes (set in env before launch)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hnm_recommendation_b200.scorer import FusedScorer
from hnm_recommendation_b200 import synth
U = int(os.environ.get("PROBE_USERS", synth.HM_USERS)); I = synth.HM_ITEMS
g = torch.Generator().manual_seed(0)
ue = (torch.randn(U, 64, generator=g) * 0.1).cuda(); ie = (torch.randn(I, 64, generator=g) * 0.1).cuda()
sc = FusedScorer(ue, ie); sc.profile = True
for rep in range(3):
    sc.topk(None, 12, fallback=False)
cnt, thr = sc._debug
print("mode", os.environ.get("HNM_FUSED_DEBUG", "0"), "users", U, "fused ms", round(sc.stage_ms["fused"], 2),
      "TFLOP/s", round(2.0 * U * I * 64 / sc.stage_ms["fused"] / 1e9, 1), "cand mean", round(float(cnt.float().sum(1).mean()), 1), flush=True)
