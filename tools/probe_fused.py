"""Dev probe: H&M-shape propagate + fused score/top-12 for all users; stage times, fallback rate, candidates."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hnm_recommendation_b200 import LightGCN, synth
from hnm_recommendation_b200.scorer import FusedScorer

kind = sys.argv[1] if len(sys.argv) > 1 else "xavier"
data = synth.interactions(synth.HM_USERS, synth.HM_ITEMS, synth.HM_EDGES)
U, I = data.num_users, data.num_items
m = LightGCN(U, I).to("cuda")
with torch.no_grad():
    w = synth.xavier_embeddings(U + I, 64) if kind == "xavier" else synth.trained_like_embeddings(U + I, 64)
    m.embeddings.weight.copy_(w)
m.set_graph(data.edge_index().cuda())
ue, ie = m.forward()
print("emb absmax", float(ue.abs().max()), float(ie.abs().max()), "item mean norm", float(ie.mean(0).norm()), "item norm mean", float(ie.norm(dim=1).mean()), flush=True)
for margin in [int(x) for x in os.environ.get('PROBE_MARGINS', '2,3,4').split(',')]:
    sc = FusedScorer(ue, ie, sel_margin=margin)
    sc.profile = True
    for rep in range(3):
        torch.cuda.synchronize(); t = time.time()
        ids, s = sc.topk(None, 12)
        torch.cuda.synchronize(); dt = time.time() - t
    cnt, thr = sc._debug
    print(f"margin {margin}: wall {dt*1e3:.1f} ms  stages {sc.stage_ms}  stats {sc.last_stats}  cand mean {float(cnt.float().sum(1).mean()):.1f} max half {int(cnt.max())}", flush=True)
    f = 2.0 * U * I * 64
    print(f"   fused TFLOP/s {f/sc.stage_ms['fused']/1e9:.1f}", flush=True)
# spot check vs exact kernel
from hnm_recommendation_b200 import engine
uids = torch.randint(0, U, (4096,), device="cuda")
e_ids, e_s = engine.topk_exact(ue, ie, uids, 12)
print("spot-check equal:", bool(torch.equal(e_ids, ids[uids])), bool(torch.equal(e_s, s[uids])))
