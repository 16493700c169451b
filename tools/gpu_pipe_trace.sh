#!/bin/bash
# Wall-clock marks of the host pipeline (DFX_HOST_PIPE_TRACE) for C2 and C5, then the e2e numbers without tracing.
timeout 120 python -m pytest tests/test_gpu_parity.py -k "host_pipeline or host_buffer or host_path or full_size_properties_c2" -x -q 2>&1 | tail -3 || exit 1
for W in c2 c5_heun; do
  DFX_HOST_PIPE_TRACE=1 timeout 100 python bench.py --workload $W --steps 3 --warmup 3 --cpu-sample 1024 2>&1 >/dev/null | grep "host pipe" | tail -2
done
for W in c2 c5_heun c5_shark; do
timeout 120 python bench.py --workload $W --steps 30 --warmup 5 --cpu-sample 4096 > gpurun_out/bench_${W}_pipe.json 2> gpurun_out/bench_${W}_pipe.err
python - <<PY
import json
for l in open("gpurun_out/bench_${W}_pipe.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print("$W", d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["ms_per_step"], d["e2e"]["value"], d["clocks"])
PY
done
