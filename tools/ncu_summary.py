"""Summarise `ncu --set full` captures for profiles/: one TSV row per captured launch, and the DRAM bytes per
launch of the dominant kernels (profiles/r2_ncu_dram_bytes.json) that bench.py reports as `roofline.traffic`
while the kernel sources are still the captured ones (their SHA-256 is stored beside the figures).

    python tools/ncu_summary.py --out profiles/r2_top_kernels.tsv [--json profiles/r2_ncu_dram_bytes.json] a.ncu-rep ...
"""
import argparse
import csv
import hashlib
import io
import json
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

COLS = [
    ("duration_ms", "gpu__time_duration.sum", 1e-6),                     # ns -> ms (unit normalised below)
    ("dram_read_GB", "dram__bytes_read.sum", 1e-9),
    ("dram_write_GB", "dram__bytes_write.sum", 1e-9),
    ("dram_pct_peak", "dram__throughput.avg.pct_of_peak_sustained_elapsed", 1),
    ("tensor_pipe_pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 1),
    ("alu_pipe_pct", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", 1),
    ("fma_pipe_pct", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", 1),
    ("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active", 1),
    ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active", 1),
    ("l2_hit_pct", "lts__t_sector_hit_rate.pct", 1),
    ("lts_throughput_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", 1),
    ("registers", "launch__registers_per_thread", 1),
    ("stall_long_sb", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", 1),
    ("stall_wait", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", 1),
    ("stall_math", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", 1),
    ("stall_branch", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", 1),
    ("stall_barrier", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", 1),
]
UNIT_SCALE = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0,
              "Tbyte": 1e3}


def rows_of(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(raw)))
    hdr, units = rd[0], rd[1]
    for r in rd[2:]:
        yield {h: (v, u) for h, u, v in zip(hdr, units, r)}


def value(row, metric, kind):
    if metric not in row:
        return None
    v, u = row[metric]
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return None
    if kind != 1:                       # a quantity with a unit: normalise to ms / GB
        return x * UNIT_SCALE.get(u, 1.0)
    return x


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("reports", nargs="+")
    ap.add_argument("--out", required=True)
    ap.add_argument("--json")
    ap.add_argument("--capture", default="")
    a = ap.parse_args()
    lines = ["\t".join(["report", "kernel"] + [c[0] for c in COLS])]
    best = {}
    for rep in a.reports:
        for row in rows_of(rep):
            name = row.get("Kernel Name", ("?", ""))[0]
            m = re.search(r"(\w+_kernel\w*|\w*Kernel\w*)", name)
            short = m.group(1) if m else name[:40]
            vals = [value(row, m, k) for _, m, k in COLS]
            lines.append("\t".join([os.path.basename(rep), short] + ["" if v is None else f"{v:.4g}" for v in vals]))
            d = dict(zip([c[0] for c in COLS], vals))
            if d["dram_read_GB"] is not None:
                best.setdefault(short, []).append(d)
    open(a.out, "w").write("\n".join(lines) + "\n")
    if a.json:
        def sha(rel):
            return hashlib.sha256(open(os.path.join(ROOT, rel), "rb").read()).hexdigest()
        kernels = {}
        fused = [d for k, v in best.items() if "score_topk_fused" in k for d in v]
        if fused:
            d = max(fused, key=lambda x: x["duration_ms"])
            kernels["fused"] = {"dram_bytes": (d["dram_read_GB"] + d["dram_write_GB"]) * 1e9,
                                "source": "hnm_recommendation_b200/csrc/score_fused.cu",
                                "sha256": sha("hnm_recommendation_b200/csrc/score_fused.cu")}
        spmm = {k: max(v, key=lambda x: x["duration_ms"]) for k, v in best.items() if k.startswith("spmm_")}
        if spmm:
            tot = sum((d["dram_read_GB"] + d["dram_write_GB"]) for d in spmm.values()) * 1e9
            kernels["spmm_layer"] = {"dram_bytes": tot, "kernels": sorted(spmm),
                                     "source": "hnm_recommendation_b200/csrc/spmm.cu",
                                     "sha256": sha("hnm_recommendation_b200/csrc/spmm.cu")}
        json.dump({"capture": a.capture, "kernels": kernels}, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
