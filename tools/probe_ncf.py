"""Dev probe: NeuralCF candidate scoring at BASELINE.json configs[3] (1.37M users x 1000 candidates)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hnm_recommendation_b200 import NeuralCF, synth

U, I, C = synth.HM_USERS, synth.HM_ITEMS, 1000
torch.manual_seed(43)
m = NeuralCF(U, I).to("cuda").eval()
chunk = 131072
g = torch.Generator(device="cuda").manual_seed(43)
cand = torch.randint(0, I, (chunk, C), dtype=torch.int32, device="cuda", generator=g)
uids = torch.arange(chunk, device="cuda")
m.score_candidates(uids, cand)               # builds the layer-1 tables
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
n = 0
for u0 in range(0, U, chunk):
    u1 = min(U, u0 + chunk)
    ids = torch.arange(u0, u1, device="cuda")
    out = m.score_candidates(ids, cand[: u1 - u0])
    n += (u1 - u0) * C
e.record(); torch.cuda.synchronize()
ms = s.elapsed_time(e)
print(f"ncf candidates: {n/1e9:.3f} G pairs in {ms:.1f} ms -> {n/ms/1e6:.2f} G pairs/s ; {n*20736/ms/1e9:.1f} TFLOP/s (reference formulation) ; compulsory bytes {n*8/ms/1e6:.0f} GB/s", flush=True)
