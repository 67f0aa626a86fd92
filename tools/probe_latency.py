"""Dev probe: recommend() latency by batch size, tensor-core path vs exact kernel (H&M shape)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hnm_recommendation_b200 import LightGCN, synth, engine
from hnm_recommendation_b200.scorer import FusedScorer

data = synth.interactions(synth.HM_USERS, synth.HM_ITEMS, synth.HM_EDGES)
U, I = data.num_users, data.num_items
m = LightGCN(U, I).to("cuda")
m.set_graph(data.edge_index().cuda())
ue, ie = m.forward()
sc = FusedScorer(ue, ie)
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.time()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.time() - t0) / n * 1e3
for B in (1, 8, 64, 256, 1024, 4096, 16384):
    uids = torch.randint(0, U, (B,), device="cuda")
    a = t(lambda: sc.topk(uids, 12))
    b = t(lambda: engine.topk_exact(ue, ie, uids, 12))
    c = t(lambda: m.recommend(uids))
    print(f"B={B:6d}  fused {a:8.3f} ms   exact {b:8.3f} ms   model.recommend {c:8.3f} ms", flush=True)
