#!/usr/bin/env python
"""Headline benchmark: users/sec, full-catalog top-12 at H&M scale (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config hm|config1|ncf]

One step = one pass of the hot path over the synthetic H&M-shaped workload
(BASELINE.json configs[1]): 3-layer dim-64 LightGCN propagation over the
1.37M x 105.5k x 31.8M graph, then exact top-12 for all 1 371 980 users
(fp16 tensor-core nomination -> fp64 rescoring -> certificate -> exact fallback).
`value` times the step with every input resident in HBM; `e2e` times the same
step through the public model API with the embedding table coming from pinned
host memory and the [users, 12] result going back to the host.

N > 1 (torchrun, one rank per GPU): users are partitioned across the ranks for the
propagation (one all-reduce of the item block per layer) and for the scoring
(hnm_recommendation_b200/dist.py; HNM_SHARD_MODE=items selects the item-catalog
sharding of BASELINE.json's north_star).  The job is the same total work at every N,
so `scaling` is "strong".

--impl reference times the CPU oracle (plain PyTorch restatement of the
reference, oracle/) on the host cores of the box; see `cpu_baseline`.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

K_TOP = 12
DIM = 64
LAYERS = 3
NCF_CANDS = 1000
METRIC = "users/sec full-catalog top-12 (LightGCN 3-layer dim-64 propagate + score + top-12)"
METRIC_NCF = "pairs/sec NeuralCF (GMF 64 + MLP [128,64,32]) candidate scoring, 1000 candidates per user"


def ncu_dram_bytes():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernels, from the committed
    `ncu --set full` capture (profiles/r2_ncu_dram_bytes.json, written by tools/ncu_summary.py together with the
    SHA-256 of the kernel sources it was taken from).  A run under a profiler is never a bench run, so the figure
    cannot be measured live; it is reported only while the sources are still the captured ones, else null."""
    import hashlib
    path = os.path.join(ROOT, "profiles", "r2_ncu_dram_bytes.json")
    if not os.path.exists(path):
        return {}, "no committed capture"
    rec = json.load(open(path))
    out = {}
    for name, k in rec.get("kernels", {}).items():
        src = os.path.join(ROOT, k["source"])
        sha = hashlib.sha256(open(src, "rb").read()).hexdigest() if os.path.exists(src) else None
        out[name] = k["dram_bytes"] if sha == k["sha256"] else None
    return out, rec.get("capture", "")


def config_dict(name, world, shard_mode="users"):
    """The workload description; identical for the GPU arm and the reference arm of the same run."""
    par = "single GPU" if world == 1 else (
        f"{shard_mode}-sharded scoring + user-partitioned propagation (item-row partial sums exchanged through "
        f"NVLink peer memory inside the layer kernels) x{world}")
    return {"workload": name, "top_k": K_TOP, "embedding_init": "xavier_uniform seed 42",
            "l2": "inputs larger than L2 (378 MB embedding table, 175 MB fp16 user operand); no explicit flush",
            "parallelism": par}

def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "tflops_burst": p["bf16_tflops"],
                "tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops_sustained": 1400.0, "source": "fallback"}


def workload(config: str):
    from hnm_recommendation_b200 import synth
    if config == "config1":
        u, i, e = synth.CONFIG1
        name = "LightGCN L3 d64, synthetic 10k users x 5k items x 200k interactions (configs[0])"
    elif config == "sweep":
        f = float(os.environ.get("HNM_SWEEP_SCALE", "1.0"))
        u, i, e = int(5_000_000 * f), int(1_000_000 * f), int(200_000_000 * f)
        name = (f"LightGCN L4 d256, synthetic {u} users x {i} items x {e} interactions (configs[4]"
                + ("" if f == 1.0 else f", scaled by {f}") + ")")
    elif config == "ncf":
        u, i, e = synth.HM_USERS, synth.HM_ITEMS, 0
        name = ("NeuralCF GMF 64 + MLP [128,64,32], 1371980 users x 1000 candidate items each out of 105542 "
                "(configs[3]); candidates distinct per user")
    else:
        u, i, e = synth.HM_USERS, synth.HM_ITEMS, synth.HM_EDGES
        name = "LightGCN L3 d64, synthetic H&M shape 1371980 users x 105542 items x 31788324 interactions (configs[1])"
    return u, i, e, name


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        # the upper half of the samples is "under load" (the sampler also sees the idle gaps)
        load = sm[len(sm) // 2:] if sm else []
        return {"sm_mhz": load[len(load) // 2] if load else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cuda_ms(fn, steps, warmup, barrier=None):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if barrier:
        barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(steps):
        fn()
    t1.record()
    torch.cuda.synchronize()
    if barrier:
        barrier()
    return t0.elapsed_time(t1) / steps


# ----------------------------------------------------------------------------------------- GPU comparators
def gpu_comparators(model, sample_users: int = 65536, chunk: int = 8192):
    """SURVEY.md 8(d) "on-box GPU comparators": the reference's own formulation in eager PyTorch on the SAME
    GPU -- torch.sparse CSR @ dense (cuSPARSE SpMM) for the propagation, chunked torch.matmul + torch.topk for
    the scoring (fp32, and with TF32 allowed).  Not the product path and outside every timed region of the
    headline; a bounded sample of users, extrapolated linearly.  Any failure is reported, never raised."""
    try:
        g = model.graph
        dev = g.rowptr.device
        n, u_total = g.num_nodes, model.num_users
        on_gpu = dev.type == "cuda"

        def ms(fn, reps):
            fn()
            if on_gpu:
                torch.cuda.synchronize(dev)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(reps):
                    fn()
                b.record()
                torch.cuda.synchronize(dev)
                return a.elapsed_time(b) / reps
            t0 = time.perf_counter()
            for _ in range(reps):
                fn()
            return (time.perf_counter() - t0) * 1e3 / reps

        with torch.no_grad():
            rowptr, col = g.rowptr.long(), g.col.long()
            row = torch.repeat_interleave(torch.arange(n, device=dev), rowptr[1:] - rowptr[:-1])
            val = g.dis[row] * g.dis[col] if g.w is None else g.dis[row] * g.w * g.dis[col]    # lightgcn.py:106
            del row
            adj = torch.sparse_csr_tensor(rowptr, col, val, size=(n, n))
            x = model.embeddings.weight.detach()
            alphas = [float(a) for a in model.alpha]

            def forward():                                         # lightgcn.py:147-158
                acc, cur = alphas[0] * x, x
                for layer in range(1, model.num_layers + 1):
                    cur = adj @ cur
                    acc = acc + alphas[layer] * cur
                return acc

            t_fwd = ms(forward, 3)
            final = forward()
            ue, ie = final[:u_total], final[u_total:]
            nu = min(sample_users, u_total)

            def score():                                           # lightgcn.py:202,356 in 8192-user batches
                for s0 in range(0, nu, chunk):
                    torch.topk(ue[s0:s0 + chunk] @ ie.t(), min(K_TOP, ie.size(0)), dim=1)

            out = {"kind": "eager PyTorch on the same GPU (torch.sparse CSR @ dense, torch.matmul + torch.topk); "
                           "the reference's formulation, not the product path",
                   "propagate_ms": t_fwd, "sample_users": nu, "chunk_users": chunk}
            prev = torch.backends.cuda.matmul.allow_tf32
            for name, flag in (("fp32", False), ("tf32", True)):
                torch.backends.cuda.matmul.allow_tf32 = flag
                t = ms(score, 2)
                out[f"score_topk_ms_{name}"] = t * u_total / nu
                out[f"users_per_s_{name}"] = u_total / (t_fwd + t * u_total / nu) * 1e3
            torch.backends.cuda.matmul.allow_tf32 = prev
            return out
    except Exception as exc:  # noqa: BLE001
        return {"error": f"{type(exc).__name__}: {exc}"[:300]}


# ----------------------------------------------------------------------------------------- CPU arm
def host_threads() -> int:
    """Use every host core for the CPU legs, whatever OMP_NUM_THREADS says (torchrun exports 1)."""
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0)) or n
    except Exception:  # noqa: BLE001
        pass
    torch.set_num_threads(n)
    return torch.get_num_threads()


class CpuArm:
    """The oracle (CPU restatement of the reference) on the host cores: graph built once, then every
    step = forward() at full shape + score/top-12 for a bounded slice of users in the reference's own
    loop shape (1024-user batches, scripts/benchmark_models.py:151-164), extrapolated linearly."""

    def __init__(self, u, i, e, seed=42):
        import oracle as O
        from hnm_recommendation_b200 import synth
        self.O, self.u = O, u
        data = synth.interactions(u, i, e, seed=seed)
        self.w = synth.xavier_embeddings(u + i, DIM, seed=seed)
        t0 = time.time()
        self.rowptr, self.col, self.val, _ = O.build_norm_adj(data.edge_index(), None, u + i)
        self.t_graph = time.time() - t0
        self.alphas = O.layer_weights(LAYERS)

    def step(self, sample_users):
        O = self.O
        n = min(sample_users, self.u)
        t0 = time.time()
        ue, ie = O.forward(self.w, self.rowptr, self.col, self.val, self.u, LAYERS, self.alphas)
        t_fwd = time.time() - t0
        t0 = time.time()
        for s0 in range(0, n, 1024):
            uids = torch.arange(s0, min(n, s0 + 1024))
            scores = O.predict_all_items(ue, ie, uids)     # scripts/benchmark_models.py:158
            torch.topk(scores, K_TOP, dim=1)               # :164
        t_score = time.time() - t0
        total = t_fwd + t_score * (self.u / n)
        return {"t_set_graph_s": self.t_graph, "t_forward_s": t_fwd, "t_score_sample_s": t_score,
                "sample_users": n, "users_per_s": self.u / total}


class NcfCpuArm:
    """oracle.ncf_forward (the reference's NeuralCF.forward, eval mode) on a bounded sample of the pairs."""

    def __init__(self, u, i, seed=43):
        import oracle as O
        torch.manual_seed(seed)
        self.O, self.u, self.i = O, u, i
        self.state = O.NeuralCFOracle(u, i).state

    def step(self, sample_pairs):
        g = torch.Generator().manual_seed(1)
        users = torch.randint(0, self.u, (sample_pairs // NCF_CANDS,), generator=g).repeat_interleave(NCF_CANDS)
        items = torch.randint(0, self.i, (users.numel(),), generator=g)
        t0 = time.time()
        with torch.no_grad():
            self.O.ncf_forward(self.state, users, items)
        dt = time.time() - t0
        return {"pairs": users.numel(), "t_s": dt, "pairs_per_s": users.numel() / dt}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_threads()
    u, i, e, name = workload(args.config)
    world = args.gpus
    cfg = config_dict(name, world, os.environ.get("HNM_SHARD_MODE", "users"))
    if args.config == "ncf":
        arm = NcfCpuArm(u, i)
        vals, walls, detail = [], [], None
        for s in range(args.warmup + args.steps):
            detail = arm.step(200_000 if s < args.warmup else 4_000_000)
            if s >= args.warmup:
                vals.append(detail["pairs_per_s"])
                walls.append(detail["t_s"])
        v = sum(vals) / len(vals)
        total_pairs = u * NCF_CANDS
        # ms_per_step is the MEASURED time of the bounded step (steps x ms_per_step is what this process spent in
        # the timed region); the whole-job figure the rate implies is reported beside it, labelled as extrapolated
        line = {"impl": "reference", "metric": METRIC_NCF, "value": v, "unit": "pairs/s", "n_gpus": args.gpus,
                "steps": len(vals), "warmup": args.warmup, "ms_per_step": 1e3 * sum(walls) / len(walls),
                "units_per_step": detail["pairs"], "full_job_ms_extrapolated": 1e3 * total_pairs / v,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": cfg,
                "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": cores, "kind": "port",
                                 "sample": f"oracle.ncf_forward on {detail['pairs']} random (user, item) pairs per step, "
                                           f"extrapolated linearly to {total_pairs} pairs"},
                "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "timing": "host wall clock, CPU only", "gpu_launches": 0}
        print(json.dumps(line))
        return
    arm = CpuArm(u, i, e)
    full = 65536 if args.config == "hm" else u          # a 65 536-user slice of configs[1]
    vals, walls, detail = [], [], None
    for s in range(args.warmup + args.steps):
        detail = arm.step(2048 if (s < args.warmup and args.config == "hm") else full)    # warm-up steps are untimed
        if s >= args.warmup:
            vals.append(detail["users_per_s"])
            walls.append(detail["t_forward_s"] + detail["t_score_sample_s"])
    v = sum(vals) / len(vals)
    sample_txt = (f"per step: oracle forward() at full shape + score/top-12 for the first {detail['sample_users']} users "
                  f"in 1024-user batches, extrapolated linearly to {u} users (graph built once: "
                  f"{detail['t_set_graph_s']:.1f} s, not counted)")
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "users/s", "n_gpus": args.gpus,
            "steps": len(vals), "warmup": args.warmup, "ms_per_step": 1e3 * sum(walls) / len(walls),
            "units_per_step": detail["sample_users"], "full_job_ms_extrapolated": 1e3 * u / v,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg,
            "cpu_baseline": {"value": v, "unit": "users/s", "cores": cores, "kind": "port", "sample": sample_txt,
                             "t_forward_s": detail["t_forward_s"], "t_score_sample_s": detail["t_score_sample_s"]},
            "e2e": {"value": v, "unit": "users/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "timing": "host wall clock, CPU only", "gpu_launches": 0}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------- NeuralCF leg
def ncf_leg(dev, steps, warmup, with_cpu=True, chunk=131072):
    """BASELINE.json configs[3]: NeuralCF logits for 1 371 980 users x 1 000 candidate items each, one GPU.
    resident: candidates already in HBM, one launch per 131 072 users; e2e: the int32 candidate ids come from
    pinned host memory and the fp32 logits go back to it, chunk by chunk on two streams."""
    from hnm_recommendation_b200 import NeuralCF, _lib, synth
    u, i, _, name = workload("ncf")
    torch.manual_seed(43)
    model = NeuralCF(u, i).to(dev).eval()
    g = torch.Generator(device=dev).manual_seed(43)
    # distinct candidates per user: one random draw of 1 000 distinct items, rotated by a per-user offset
    base = torch.randperm(i, device=dev, generator=g)[:NCF_CANDS].to(torch.int32)
    off = torch.randint(0, i, (u, 1), device=dev, generator=g, dtype=torch.int32)
    cand = (off + base.unsqueeze(0)) % i                                   # [u, 1000] int32, 5.5 GB
    del off
    out = torch.empty(u, NCF_CANDS, dtype=torch.float32, device=dev)       # 5.5 GB
    ranges = [(a, min(u, a + chunk)) for a in range(0, u, chunk)]

    def step_resident():
        for a, b in ranges:
            out[a:b] = model.score_candidates(torch.arange(a, b, device=dev), cand[a:b])

    model.score_candidates(torch.arange(0, 8, device=dev), cand[:8])       # builds the layer-1 tables
    _lib.LAUNCHES = 0
    ms = cuda_ms(step_resident, steps, warmup)
    launches = _lib.LAUNCHES // max(1, steps + warmup)
    pairs = u * NCF_CANDS
    # parity spot check against the oracle on the same parameters (outside the timed region)
    check = None
    try:
        import oracle as O
        state = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        rows = torch.tensor([0, 1, u // 2, u - 1])
        want = O.ncf_forward(state, rows.repeat_interleave(NCF_CANDS), cand[rows.to(dev)].cpu().long().view(-1))
        got = out[rows.to(dev)].cpu().view(-1)
        scale = float(want.abs().max())
        check = {"pairs": int(want.numel()), "max_abs_err": float((got - want).abs().max()),
                 "tolerance": f"rtol 1e-5 + atol 1e-6*max = {1e-6 * scale:.2e}",
                 "ok": bool(((got - want).abs() <= 1e-5 * want.abs() + 1e-6 * scale).all())}
    except Exception as exc:  # noqa: BLE001
        check = {"error": f"{type(exc).__name__}: {exc}"[:200]}
    # e2e: host ids in, host logits out
    e2e = None
    try:
        n_e2e = min(u, 4 * chunk)                        # a bounded slice keeps the pinned buffers at 2 x 2.1 GB
        cand_host = cand[:n_e2e].cpu().pin_memory()
        out_host = torch.empty(n_e2e, NCF_CANDS, dtype=torch.float32).pin_memory()
        copy = torch.cuda.Stream(device=dev)
        bufs = [torch.empty(chunk, NCF_CANDS, dtype=torch.int32, device=dev) for _ in range(2)]

        def step_e2e():
            cur = torch.cuda.current_stream(dev)
            for n, a in enumerate(range(0, n_e2e, chunk)):
                b = min(n_e2e, a + chunk)
                buf = bufs[n & 1]
                buf[: b - a].copy_(cand_host[a:b], non_blocking=True)
                res = model.score_candidates(torch.arange(a, b, device=dev), buf[: b - a])
                copy.wait_stream(cur)
                with torch.cuda.stream(copy):
                    out_host[a:b].copy_(res, non_blocking=True)
                    res.record_stream(copy)
                cur.wait_stream(copy) if n & 1 else None
            copy.synchronize()
            cur.synchronize()

        ms_e2e = cuda_ms(step_e2e, max(2, steps // 2), 1)
        e2e = {"value": n_e2e * NCF_CANDS / ms_e2e * 1e3, "unit": "pairs/s", "ms_per_step": ms_e2e,
               "pairs_per_step": n_e2e * NCF_CANDS, "h2d_bytes_per_step": n_e2e * NCF_CANDS * 4,
               "d2h_bytes_per_step": n_e2e * NCF_CANDS * 4,
               "note": f"first {n_e2e} users (bounded pinned buffers); the rate does not depend on the user count"}
        del cand_host, out_host, bufs
    except Exception as exc:  # noqa: BLE001
        e2e = {"error": f"{type(exc).__name__}: {exc}"[:200]}
    pk = peaks()
    tables = (2 * u + 2 * i) * 64 * 4
    compulsory = pairs * 8 + tables                      # ids in + logits out + every table row once
    res = {"metric": METRIC_NCF, "value": pairs / ms * 1e3, "unit": "pairs/s", "ms_per_step": ms,
           "config": {"workload": name}, "gpu_launches": launches, "parity_check": check, "e2e": e2e,
           "roofline": {"bound": "hbm", "kernel": "ncf_score_tc_kernel", "achieved": compulsory / ms / 1e6,
                        "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": compulsory / ms / 1e6 / pk["hbm_gbs"],
                        "algorithmic_bytes": compulsory, "traffic": None,
                        "gather_model_gbs": pairs * 520 / ms / 1e6,
                        "note": "compulsory bytes = ids + logits + tables once; the kernel's own bound is the "
                                "L2 gather of 512 B of item rows per pair (gather_model_gbs)"},
           "tflops_reference_formulation": pairs * 20736 / ms / 1e9}
    if with_cpu:
        cores = host_threads()
        c = NcfCpuArm(u, i).step(4_000_000)
        res["cpu_baseline"] = {"value": c["pairs_per_s"], "unit": "pairs/s", "cores": cores, "kind": "port",
                               "sample": f"oracle.ncf_forward on {c['pairs']} random pairs ({c['t_s']:.2f} s)"}
    del cand, out, model
    torch.cuda.empty_cache()
    return res


def run_ncf(args):
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    sampler = ClockSampler(0)
    sampler.start()
    r = ncf_leg(dev, args.steps, args.warmup, with_cpu=not args.no_cpu)
    clocks = sampler.stop()
    line = {"metric": r["metric"], "value": r["value"], "unit": r["unit"], "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32 (layer 2: 3xTF32 tensor-core split, fp32 accumulate)",
            "data": "synthetic", "config": r["config"], "e2e": r["e2e"], "gpu_launches": r["gpu_launches"],
            "clocks": clocks, "roofline": r["roofline"], "parity_check": r["parity_check"]}
    if "cpu_baseline" in r:
        line["cpu_baseline"] = r["cpu_baseline"]
    print(json.dumps(line))


def bind_to_gpu_numa(gpu_index: int):
    """Run this rank on the CPUs of its GPU's NUMA node BEFORE the pinned host buffers are allocated (first touch
    puts them on that node): with eight ranks' buffers all on node 0 the host-to-device copies of the e2e leg
    shared one socket's memory controllers.  Best effort; returns the node or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        path = f"/sys/bus/pci/devices/{bus[-12:].lower()}/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.extend(range(int(a), int(b or a) + 1))
        allowed = set(os.sched_getaffinity(0)) & set(cpus)
        if allowed:
            os.sched_setaffinity(0, allowed)
        return node
    except Exception:  # noqa: BLE001
        return None


# ----------------------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    import torch.distributed as dist
    from hnm_recommendation_b200 import LightGCN, _lib, engine, synth
    from hnm_recommendation_b200 import dist as hdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa(local) if world > 1 else None
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")     # NCCL's version banner goes to stdout; the line below is the output
        dist.init_process_group("nccl", device_id=dev)
    barrier = (lambda: dist.barrier()) if world > 1 else None

    u, i, e, name = workload(args.config)
    sweep = args.config == "sweep"
    DIM, LAYERS = (256, 4) if sweep else (globals()["DIM"], globals()["LAYERS"])
    model = LightGCN(u, i, embedding_dim=DIM, num_layers=LAYERS, top_k=K_TOP).to(dev)
    if sweep:
        # configs[4]: the graph is drawn once (rank 0) and broadcast; the 6 M x 256 table is initialised on the
        # device (N(0, 0.1), same seed on every rank); no host copy of the table -> no e2e leg for this config
        m2 = 2 * e
        ei = torch.empty(2, m2, dtype=torch.int64, device=dev)
        if rank == 0:
            ei.copy_(synth.interactions(u, i, e, seed=42).edge_index())
        if world > 1:
            dist.broadcast(ei, src=0)
        with torch.no_grad():
            model.embeddings.weight.normal_(0.0, 0.1, generator=torch.Generator(device=dev).manual_seed(42))
        model.set_graph(ei)
        del ei
        w_host = out_host = None
    else:
        data = synth.interactions(u, i, e, seed=42)
        w_host = synth.xavier_embeddings(u + i, DIM, seed=42).pin_memory()
        with torch.no_grad():
            model.embeddings.weight.copy_(w_host)
        model.set_graph(data.edge_index().to(dev))
        del data
        out_host = torch.empty(u, K_TOP, dtype=torch.int64).pin_memory()
    model.cache_embeddings = False                  # every step recomputes the propagation
    sharded = hdist.ShardedLightGCN(model) if world > 1 else None

    def step_resident():
        return sharded.recommend_all() if sharded else model.recommend_all()

    if sharded and sharded.mode == "users":
        # the table crosses PCIe once over the job: a rank uploads its own users' rows and its 1/G slice of the item
        # block (completed over NVLink: ShardedLightGCN.load_embeddings_from_host), and owns one slice of the result
        u0, u1 = sharded.plan.user_rows[rank]
        h2d_rows = [(u0, u1), sharded.plan.item_rows[rank]]
        d2h_rows = (u0, u1)
    elif sharded:
        h2d_rows = [sharded.plan.user_rows[rank], sharded.plan.item_rows[rank]]
        d2h_rows = (0, u) if rank == 0 else (0, 0)
    else:
        h2d_rows = [(0, u + i)]
        d2h_rows = (0, u)
    h2d_bytes = sum(b - a for a, b in h2d_rows) * DIM * 4
    d2h_bytes = (d2h_rows[1] - d2h_rows[0]) * K_TOP * 8

    def step_e2e():
        if sharded:
            sharded.load_embeddings_from_host(w_host)                       # H2D of the step's input (+ NVLink)
            ids = sharded.recommend_all()
            a, b = d2h_rows
            out_host[a:b].copy_(ids[a:b], non_blocking=True)                # D2H of the step's result
        else:
            model.embeddings.weight.data.copy_(w_host, non_blocking=True)   # H2D of the step's input
            model.recommend_all(out_host=out_host)                          # D2H streamed chunk by chunk
        torch.cuda.current_stream().synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    _lib.LAUNCHES = 0
    ms = cuda_ms(step_resident, args.steps, args.warmup, barrier)
    launches = _lib.LAUNCHES // max(1, args.steps + args.warmup)
    ms_e2e = cuda_ms(step_e2e, max(2, args.steps // 2), 1, barrier) if not sweep else float("nan")
    clocks = sampler.stop() if rank == 0 else None
    # the upload alone (same copies as in step_e2e), to split the e2e overhead into PCIe and the rest
    h2d_ms = float("nan")
    if not sweep:
        def upload():
            if sharded:
                sharded.load_embeddings_from_host(w_host)
            else:
                model.embeddings.weight.data.copy_(w_host, non_blocking=True)
        h2d_ms = cuda_ms(upload, 3, 1, barrier)
    if world > 1:
        t = torch.tensor([ms, ms_e2e, h2d_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e, h2d_ms = t.tolist()
        t = torch.tensor([h2d_bytes, d2h_bytes], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        h2d_bytes, d2h_bytes = (int(x) for x in t.tolist())

    # correctness of the multi-GPU data path, on record with every line: 2 048 random users of the all-gathered
    # result against the brute-force kernel on the all-gathered embeddings of the same sharded propagation
    mg_check = None
    items_mode_ms = None
    if sharded is not None:
        with torch.no_grad():
            ids = sharded.recommend_all()
            ue, ie = sharded.forward(all_rows=True)
            uids = torch.randint(0, u, (2048,), device=dev, generator=torch.Generator(device=dev).manual_seed(11))
            w_ids, _ = engine.topk_exact(ue.contiguous(), ie.contiguous(), uids, K_TOP)
            bad = int((ids[uids] != w_ids).any(dim=1).sum())
        t = torch.tensor([bad], device=dev, dtype=torch.int64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        mg_check = {"users": 2048, "mismatches": int(t.item()),
                    "against": "hnm_topk_exact on the all-gathered embeddings of the sharded propagation, every rank"}
        # north_star's own partitioning (item-catalog shards + all-to-all + hnm_merge_topk), same job
        try:
            if sweep:
                raise RuntimeError("not run for configs[4] (every rank would hold candidate lists for all 5 M users)")
            sh_items = hdist.ShardedLightGCN(model, mode="items")
            items_mode_ms = cuda_ms(lambda: sh_items.recommend_all(), 2, 1, barrier)
            t = torch.tensor([items_mode_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            items_mode_ms = float(t.item())
            del sh_items
        except Exception as exc:  # noqa: BLE001
            items_mode_ms = f"{type(exc).__name__}: {exc}"[:200]

    # per-stage kernel times on this rank (CUDA events on the launching stream)
    stages = hdist.profile_stages(model, sharded, steps=max(2, args.steps // 2))
    if sharded is None and not sweep:
        # the serving default of the reference (scripts/serve.py:350-352): every user's own purchases filtered out
        stages["filtered_step_ms"] = cuda_ms(lambda: model.recommend_all(filter_purchased=True), 2, 1, None)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    nnz = 2 * e + (u + i)
    n_nodes = u + i
    flops = 2.0 * u * i * DIM
    spmm_alg_bytes = 2 * n_nodes * DIM * 4 + nnz * 4 + (n_nodes + 1) * 4 + n_nodes * 4
    fused_ms = stages["fused_ms"]
    ach_tf = flops / world / fused_ms / 1e9 if fused_ms else None
    # t_score_topk includes rescoring, merge and fallback
    topk_ms = (stages["fused_ms"] + stages["rescore_ms"] + stages["fallback_ms"] + stages["pack_users_ms"]) or None
    ach_topk = flops / world / topk_ms / 1e9 if topk_ms else None
    spmm_ms = stages["spmm_layer_ms"]
    prop_ms = stages.get("propagate_ms")
    traffic, traffic_src = ncu_dram_bytes()
    hm1 = world == 1 and args.config == "hm"
    line = {
        "metric": METRIC if not sweep else METRIC.replace("3-layer dim-64", "4-layer dim-256"),
        "value": u / ms * 1e3, "unit": "users/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32 propagate / f16xf16->f32 tensor-core nomination / f64 exact rescoring",
        "data": "synthetic",
        "config": config_dict(name, world, sharded.mode if sharded else "users"),
        "e2e": ({"value": u / ms_e2e * 1e3, "unit": "users/s", "ms_per_step": ms_e2e,
                 "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": int(d2h_bytes),
                 "h2d_ms_alone": h2d_ms,
                 "how": ("every rank uploads its own users' rows and its 1/G slice of the item block from pinned host "
                         "memory (the table crosses PCIe once over the job), the item block is completed over NVLink; "
                         "every rank reads its own slice of the all-gathered result back" if world > 1 else
                         "whole table uploaded from pinned host memory, result streamed back chunk by chunk")}
                if not sweep else
                {"value": None, "unit": "users/s", "note": "not measured for configs[4]: the table is initialised on the device"}),
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "score_topk_fused_kernel", "achieved": ach_tf,
                     "peak": pk["tflops_burst"], "unit": "TFLOP/s",
                     "frac": ach_tf / pk["tflops_burst"] if ach_tf else None,
                     "frac_of_sustained_peak": ach_tf / pk["tflops_sustained"] if ach_tf else None,
                     "score_topk": {"ms": topk_ms, "achieved": ach_topk,
                                    "frac": ach_topk / pk["tflops_burst"] if ach_topk else None,
                                    "what": "pack + fused select + exact rescoring + fallback tiers (SURVEY 8d t_score_topk)"},
                     "traffic": traffic.get("fused") if hm1 else None,
                     "traffic_source": traffic_src if hm1 else "single-GPU configs[1] capture only",
                     "peak_source": pk["source"] + " (cuBLAS bf16 burst; sustained %.1f)" % pk["tflops_sustained"],
                     "algorithmic_flops": flops / world, "ms": fused_ms},
        "roofline_spmm": {"bound": "hbm", "kernel": "spmm layer kernels (one layer)",
                          "achieved": spmm_alg_bytes / world / spmm_ms / 1e6 if spmm_ms else None,
                          "peak": pk["hbm_gbs"], "unit": "GB/s",
                          "frac": spmm_alg_bytes / world / spmm_ms / 1e6 / pk["hbm_gbs"] if spmm_ms else None,
                          "gather_model_gbs": (nnz * (4 + 4 * DIM) + n_nodes * 4 * DIM) / world / spmm_ms / 1e6 if spmm_ms else None,
                          "gather_ceiling_gbs": {"l2_resident_27MB_table": 19320.0, "dram_351MB_table": 8940.0,
                                                 "source": "tools/bench_l2_gather.cu on this pool's B200 "
                                                           "(profiles/r2_l2_gather_microbench.txt)"},
                          "algorithmic_bytes": spmm_alg_bytes / world, "ms": spmm_ms,
                          # SURVEY 8(d)'s own definition: B_alg * L over the whole propagation (prescale, all layers,
                          # and at N > 1 the exchanges) as forward() runs it
                          "propagate": ({"ms": prop_ms, "layers": LAYERS,
                                         "achieved": spmm_alg_bytes * LAYERS / world / prop_ms / 1e6,
                                         "frac": spmm_alg_bytes * LAYERS / world / prop_ms / 1e6 / pk["hbm_gbs"]}
                                        if prop_ms else None),
                          "traffic": traffic.get("spmm_layer") if hm1 else None},
        "stages_ms": stages,
    }
    if mg_check is not None:
        line["multi_gpu_check"] = mg_check
        line["items_mode_ms"] = items_mode_ms
        line["host_numa_node_rank0"] = numa
    if sweep:
        line["config"]["embedding_init"] = "normal(0, 0.1) on the device, seed 42"
        line["config"]["l2"] = "inputs far larger than L2 (6.1 GB table)"
    if world == 1 and not args.no_cpu and not sweep:
        line["gpu_comparators"] = gpu_comparators(model)
        cores = host_threads()
        cpu = CpuArm(u, i, e).step(65536 if args.config == "hm" else u)    # a 65 536-user slice
        line["cpu_baseline"] = {
            "value": cpu["users_per_s"], "unit": "users/s", "cores": cores, "kind": "port",
            "sample": (f"oracle forward() at full shape ({cpu['t_forward_s']:.2f} s) + score/top-12 for the first "
                       f"{cpu['sample_users']} users in 1024-user batches ({cpu['t_score_sample_s']:.2f} s), "
                       f"extrapolated linearly to {u} users"),
            "t_forward_s": cpu["t_forward_s"], "t_score_sample_s": cpu["t_score_sample_s"]}
    if world == 1 and args.config == "hm" and not args.no_ncf and not sweep:
        del model, out_host
        torch.cuda.empty_cache()
        try:
            line["ncf"] = ncf_leg(dev, max(2, args.steps // 2), 3, with_cpu=not args.no_cpu)
        except Exception as exc:  # noqa: BLE001
            line["ncf"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="hm", choices=["hm", "config1", "ncf", "sweep"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-ncf", action="store_true", help="skip the NeuralCF (configs[3]) object of the headline line")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0 if args.impl == "reference" else 3)
    if args.impl == "reference":
        run_reference(args)
    elif args.config == "ncf":
        run_ncf(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
