#!/bin/bash
# Host-pipelined path: its parity tests first (short timeout: a hang must not hold the box), then the whole GPU suite and C2 / C5 bench lines.
timeout 120 python -m pytest tests/test_gpu_parity.py -k "host_pipeline or host_buffer or host_path" -x -q 2>&1 | tail -5 || exit 1
timeout 200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for W in c2 c5_heun; do
timeout 120 python bench.py --workload $W --steps 30 --warmup 5 --cpu-sample 4096 > gpurun_out/bench_${W}_pipe.json 2> gpurun_out/bench_${W}_pipe.err
python - <<PY
import json
for l in open("gpurun_out/bench_${W}_pipe.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print("$W", d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["ms_per_step"], d["e2e"]["value"], d["clocks"])
PY
done
for C in 4 8 16 32; do DFX_HOST_CHUNKS=$C timeout 100 python bench.py --workload c2 --steps 20 --warmup 3 --cpu-sample 1024 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('chunks $C e2e ms', d['e2e']['ms_per_step'])
"; done
