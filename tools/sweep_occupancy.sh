#!/bin/bash
# A/B of register / occupancy targets for the C2 kernel (rebuilds only the Lorenz TU on the GPU box).
for mb in 1 6 8; do
  export DFX_NVCC_EXTRA="-DDFX_MIN_BLOCKS=$mb"
  touch diffrax_b200/csrc/inst_lorenz.cu
  python - <<PY
import os, subprocess, glob
from diffrax_b200 import build as b
src=os.path.join(b.CSRC,'inst_lorenz.cu'); print(b._compile(src)[2:4])
objs=[os.path.join(b.OBJ, os.path.basename(s)[:-3]+'.o') for s in sorted(glob.glob(os.path.join(b.CSRC,'*.cu')))]
subprocess.run([b.NVCC,*b.ARCH,'-shared','-o',b.LIB,*objs,'-lcudart'],check=True)
PY
  grep -A3 "LorenzENS_6Dopri5ELi0ELb0" diffrax_b200/csrc/_obj/inst_lorenz.o.log | grep -E "registers|spill" | head -2
  python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('MINBLOCKS=$mb', d['ms_per_step'], d['roofline']['frac'])"
done
