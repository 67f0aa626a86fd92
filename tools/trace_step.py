"""Dev probe: kernel timeline of the sharded step (torch.profiler / CUPTI), rank 0: where the GPU idles.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/trace_step.py
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hnm_recommendation_b200 import LightGCN, synth              # noqa: E402
from hnm_recommendation_b200 import dist as hdist                # noqa: E402


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=dev)
    u, i, e = synth.HM_USERS, synth.HM_ITEMS, synth.HM_EDGES
    scale = float(os.environ.get("TRACE_USER_SCALE", "1"))       # < 1: fewer users per rank (the 8-GPU shape on 2 GPUs)
    u = int(u * scale)
    e = int(e * scale)
    data = synth.interactions(u, i, e, seed=42)
    model = LightGCN(u, i, embedding_dim=64, num_layers=3, top_k=12).to(dev)
    with torch.no_grad():
        model.embeddings.weight.copy_(synth.xavier_embeddings(u + i, 64, seed=42))
    model.set_graph(data.edge_index().to(dev))
    model.cache_embeddings = False
    sharded = hdist.ShardedLightGCN(model) if world > 1 else None
    step = (lambda: sharded.recommend_all()) if sharded else (lambda: model.recommend_all())
    for _ in range(4):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    from torch.profiler import ProfilerActivity, profile
    steps = 4
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(steps):
            step()
        torch.cuda.synchronize()
    if rank == 0:
        ev = [x for x in prof.events() if x.device_type == torch.autograd.DeviceType.CUDA]
        ev.sort(key=lambda x: x.time_range.start)
        t0, t1 = ev[0].time_range.start, max(x.time_range.end for x in ev)
        busy, gaps, cur_end, prev = 0.0, [], ev[0].time_range.start, None
        for x in ev:
            s, en = x.time_range.start, x.time_range.end
            if s > cur_end:
                gaps.append((s - cur_end, prev.name[:60] if prev else "", x.name[:60]))
                busy += en - s
                cur_end = en
            elif en > cur_end:
                busy += en - cur_end
                cur_end = en
            prev = x
        agg = {}
        for x in ev:
            k = x.name[:70]
            a = agg.setdefault(k, [0, 0.0])
            a[0] += 1
            a[1] += x.time_range.end - x.time_range.start
        out = {"world": world, "users": u, "steps": steps, "span_ms_per_step": (t1 - t0) / 1e3 / steps,
               "busy_ms_per_step": busy / 1e3 / steps, "idle_ms_per_step": (t1 - t0 - busy) / 1e3 / steps,
               "kernels_ms_per_step": {k: [v[0] / steps, v[1] / 1e3 / steps] for k, v in
                                       sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]},
               "largest_gaps_us": [(round(g, 1), a, b) for g, a, b in sorted(gaps, reverse=True)[:25]],
               "gap_count": len(gaps)}
        print(json.dumps(out, indent=1))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
