#!/bin/bash
# usage: tools/sweep_env.sh VAR "v1 v2 ..." [bench args]   -- ms_per_step of bench.py under each value of an env var
var=$1; vals=$2; shift 2
for v in $vals; do
  out=$(env $var=$v timeout 200 python bench.py --steps 10 --warmup 3 --cpu-sample 4096 "$@" 2>&1 | tail -1)
  echo "$var=$v $(echo "$out" | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"],4), "frac", round(d["roofline"]["frac"],4))' 2>&1 | tail -1)"
done
