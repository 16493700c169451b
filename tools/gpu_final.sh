#!/bin/bash
# Round-end check: the whole GPU suite, smoke(), and the default bench line (as the driver runs it).
timeout 200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 60 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 200 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
tail -c 600 gpurun_out/bench_default.json
