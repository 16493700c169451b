#!/bin/bash
# Re-measure every workload and re-profile the headline kernels; outputs land in gpurun_out/ (copy into profiles/ by hand).
set -u
tag=${1:-r01}
for w in c2 c1 c3 c4 c5_heun c5_shark; do
  timeout 400 python bench.py --workload $w 2>/dev/null | tail -1 > gpurun_out/${tag}_bench_$w.json
done
timeout 200 python bench.py --workload published_jump_step 2>/dev/null | tail -1 > gpurun_out/${tag}_bench_published_jump_step.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/${tag}_bench_reference_arm.json
# launch lists (cold-cache, serialised: shares only)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_c2_launches.csv python bench.py --steps 2 --warmup 1 --cpu-sample 1024 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_c3_launches.csv python bench.py --workload c3 --steps 2 --warmup 1 --cpu-sample 64 > /dev/null 2>&1
# full captures of the dominant kernels
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ensemble_kernel -c 1 -s 2 -f -o gpurun_out/${tag}_c2_full python bench.py --steps 2 --warmup 1 --cpu-sample 1024 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ensemble_kernel -c 1 -s 2 -f -o gpurun_out/${tag}_c3_full python bench.py --workload c3 --steps 2 --warmup 1 --cpu-sample 64 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ensemble_kernel -c 1 -s 2 -f -o gpurun_out/${tag}_c5_heun_full python bench.py --workload c5_heun --steps 2 --warmup 1 --cpu-sample 64 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_tc -c 1 -s 2 -f -o gpurun_out/${tag}_c4_full python bench.py --workload c4 --steps 2 --warmup 1 --cpu-sample 64 > /dev/null 2>&1
# keep only the text summaries (the .ncu-rep files are tens of MB each; gpurun_out/ is capped at 64 MiB)
for k in c2 c3 c5_heun c4; do
  python tools/ncu_summary.py gpurun_out/${tag}_${k}_full.ncu-rep > gpurun_out/${tag}_${k}_ncu_full.txt 2>/dev/null
  rm -f gpurun_out/${tag}_${k}_full.ncu-rep
done
ls -la gpurun_out | tail -24
