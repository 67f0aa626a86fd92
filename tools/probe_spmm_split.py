"""Dev probe: time one LightGCN layer over the user rows and over the item rows separately (H&M shape)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hnm_recommendation_b200 import LightGCN, synth
from hnm_recommendation_b200._lib import call, ptr, stream

def ev_time(fn, n=10, warm=3):
    for _ in range(warm): fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n

data = synth.interactions(synth.HM_USERS, synth.HM_ITEMS, synth.HM_EDGES)
U, I = data.num_users, data.num_items
m = LightGCN(U, I).to("cuda")
m.set_graph(data.edge_index().cuda())
g = m.graph
w = m.embeddings.weight.detach()
n, d = w.shape
xs, xo, acc = torch.empty_like(w), torch.empty_like(w), torch.empty_like(w)
call("hnm_lightgcn_prescale", ptr(w), ptr(g.dis), 0.25, ptr(xs), ptr(acc), n, d, stream())
heavy = ptr(g.heavy_rows) if g.num_heavy else None

def layer(r0, r1, with_heavy=True, short=0):
    call("hnm_lightgcn_layer", ptr(g.rowptr), ptr(g.col), ptr(g.w), ptr(g.dis), ptr(xs), ptr(xo), ptr(acc), 0.25, n, d,
         r0, r1, heavy if with_heavy else None, g.num_heavy if with_heavy else 0, g.num_huge if with_heavy else 0,
         g.heavy_threshold, short, stream())

print("all rows      %.3f ms" % ev_time(lambda: layer(0, n)))
print("user rows     %.3f ms" % ev_time(lambda: layer(0, U)))
print("user rows, staged 8-row kernel  %.3f ms" % ev_time(lambda: layer(0, U, short=1)))
print("user rows, staged, no heavy launches  %.3f ms" % ev_time(lambda: layer(0, U, with_heavy=False, short=1)))
print("item rows     %.3f ms" % ev_time(lambda: layer(U, n)))
print("item rows, no heavy kernels (heavy rows skipped) %.3f ms" % ev_time(lambda: layer(U, n, True) if False else layer(U, n)))
deg = (g.rowptr[1:] - g.rowptr[:-1]).long()
print("user deg mean %.1f max %d ; item deg mean %.1f max %d ; heavy %d huge %d" % (
    deg[:U].float().mean(), deg[:U].max(), deg[U:].float().mean(), deg[U:].max(), g.num_heavy, g.num_huge))
