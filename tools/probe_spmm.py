"""Dev probe: time set_graph / propagate / exact top-k at the H&M shape on one GPU."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hnm_recommendation_b200 import LightGCN, synth, engine

def ev_time(fn, n=5, warm=2):
    for _ in range(warm): fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n

shape = (synth.HM_USERS, synth.HM_ITEMS, synth.HM_EDGES) if len(sys.argv) < 2 else synth.CONFIG1
t = time.time(); data = synth.interactions(*shape); print("synth s", time.time() - t, flush=True)
U, I = data.num_users, data.num_items
m = LightGCN(U, I).to("cuda")
ei = data.edge_index().cuda()
torch.cuda.synchronize(); t = time.time(); m.set_graph(ei); torch.cuda.synchronize(); print("set_graph s", time.time() - t, "heavy", m.graph.num_heavy, flush=True)
del ei
m.cache_embeddings = False
ms = ev_time(lambda: m.forward())
nnz = m.graph.nnz; N = U + I
balg = (2 * N * 64 * 4 + nnz * 4 + (N + 1) * 4 + N * 4) * 3
print(f"forward {ms:.3f} ms  ({ms/3:.3f} ms/layer incl prescale)  B_alg GB/s {balg/ms/1e6:.1f}  gather GB/s {(nnz*260*3)/ms/1e6:.1f}", flush=True)
ue, ie = m.forward()
uids = torch.arange(0, 8192, device="cuda")
ms = ev_time(lambda: engine.topk_exact(ue, ie, uids, 12), n=2, warm=1)
print(f"topk_exact 8192 users {ms:.2f} ms -> {8192/ms*1e3:.0f} users/s", flush=True)
ms = ev_time(lambda: engine.score_all_items(ue, ie, uids[:1024]), n=3, warm=1)
print(f"score_all_items 1024 users {ms:.2f} ms", flush=True)
