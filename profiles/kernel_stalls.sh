#!/bin/bash
# Per-instruction stall profile of one kernel of the two-kernel gradient (1 M individuals x 2 starts):
#   [SKIP=n] bash profiles/kernel_stalls.sh <kernel-regex> <name>   -> gpurun_out/r02_<name>_stalls_{sass,cuda}.txt + ncu summary
set -e
K=$1; NAME=$2
cat > /tmp/once.py <<'PY'
import os, sys
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import bench, conditional_ude_b200 as cu
ctx = cu.Context(0)
n, S = 1_000_000, 2
pop = cu.Population(packed=bench.synthetic_population(n, 1000, bench.simulate_gpu(ctx)), ctx=ctx)
neural, cond = bench.synthetic_starts(n, S, 11, 2000)
pop.loss_grad(neural, cond, mean=False)
print(ctx.stats())
PY
python /tmp/once.py
ncu --set full --clock-control none --import-source on -k regex:$K -s ${SKIP:-0} -c 1 -o /tmp/$NAME python /tmp/once.py > /tmp/ncu.log 2>&1 || tail -5 /tmp/ncu.log
ncu -i /tmp/$NAME.ncu-rep --page source --csv > /tmp/${NAME}_sass.csv
ncu -i /tmp/$NAME.ncu-rep --page source --csv --print-source cuda > /tmp/${NAME}_cuda.csv || true
ncu -i /tmp/$NAME.ncu-rep --page raw --csv > /tmp/${NAME}_raw.csv
python profiles/summarize_ncu.py /tmp/${NAME}_raw.csv > gpurun_out/r02_${NAME}_ncu_summary.txt
python profiles/stall_by_line.py /tmp/${NAME}_sass.csv 40 > gpurun_out/r02_${NAME}_stalls_sass.txt
python profiles/stall_by_line.py /tmp/${NAME}_cuda.csv 60 > gpurun_out/r02_${NAME}_stalls_cuda.txt || true
