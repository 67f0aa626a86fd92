"""Dev probe: the fused scorer at the user counts a rank sees at 1 / 2 / 4 / 8 GPUs (H&M shape), one GPU."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hnm_recommendation_b200 import LightGCN, synth
from hnm_recommendation_b200.scorer import FusedScorer

data = synth.interactions(synth.HM_USERS, synth.HM_ITEMS, synth.HM_EDGES)
U, I = data.num_users, data.num_items
m = LightGCN(U, I).to("cuda")
with torch.no_grad():
    m.embeddings.weight.copy_(synth.xavier_embeddings(U + I, 64))
m.set_graph(data.edge_index().cuda())
ue, ie = m.forward()


def ms(fn, reps=5):
    fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for g in (1, 2, 4, 8):
    n = -(-U // g)
    sub = ue[:n]
    build = ms(lambda: FusedScorer(sub, ie))
    sc = FusedScorer(sub, ie)
    whole = ms(lambda: sc.topk(None, 12))
    sc.profile = True
    sc.topk(None, 12)
    sc.topk(None, 12)
    print(f"G={g} users={n}: scorer set-up {build:.3f} ms, topk {whole:.3f} ms (x{g} = {whole * g:.2f}), stages {sc.stage_ms}", flush=True)
