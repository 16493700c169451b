#!/bin/bash
# One short GPU pass: parity tests, smoke, a C2 bench line (printed compactly).
timeout 150 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 60 python __graft_entry__.py smoke 2>&1 | tail -2
W=${1:-c2}
timeout 120 python bench.py --workload $W --steps 30 --warmup 5 --cpu-sample 4096 > gpurun_out/bench_${W}_check.json 2> gpurun_out/bench_${W}_check.err
python - <<PY
import json
for l in open("gpurun_out/bench_${W}_check.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["clocks"])
PY
