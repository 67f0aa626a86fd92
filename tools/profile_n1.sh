#!/bin/bash
# Round-end single-GPU record: the default bench line, the ncu launch list of the same program, and one
# `--set full` capture of the hot kernels.  Usage (on the GPU box): tools/profile_n1.sh <outdir>
out=${1:-gpurun_out/r2z}
mkdir -p "$out"
timeout 600 python bench.py > "$out/bench_n1.json" 2> "$out/bench_n1.err"; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > "$out/bench_ref.json" 2> "$out/bench_ref.err"; echo "ref rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file "$out/launches_bench_n1.csv" python bench.py --steps 2 --warmup 3 --no-cpu --no-ncf > "$out/ncu_launch.log" 2>&1
echo "launch list rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none \
  -k 'regex:score_topk_fused|merge_split|rescore|spmm_|prescale|pack_kernel|colsum|absmax' --launch-skip 44 -c 24 \
  -f -o "$out/top_kernels" python bench.py --steps 2 --warmup 3 --no-cpu --no-ncf > "$out/ncu_full.log" 2>&1
echo "full capture rc=$?"
ls -la "$out"
