#!/bin/bash
# The `--set full` capture of tools/profile_n1.sh alone (after a kernel source changed), plus the launch list.
out=${1:-gpurun_out/r2z}
mkdir -p "$out"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file "$out/launches_bench_n1.csv" python bench.py --steps 2 --warmup 3 --no-cpu --no-ncf > "$out/ncu_launch.log" 2>&1
echo "launch list rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none \
  -k 'regex:score_topk_fused|merge_split|rescore|spmm_|prescale|pack_kernel|colsum|colmean|absmax' --launch-skip 46 -c 25 \
  -f -o "$out/top_kernels" python bench.py --steps 2 --warmup 3 --no-cpu --no-ncf > "$out/ncu_full.log" 2>&1
echo "full capture rc=$?"
