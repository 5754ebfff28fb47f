set -e
cat > /tmp/adj_once.py <<'PY'
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
python /tmp/adj_once.py
ncu --set full --clock-control none --import-source on -k regex:cude_adjoint_kernel -c 1 -o /tmp/adj python /tmp/adj_once.py > /tmp/ncu.log 2>&1 || tail -5 /tmp/ncu.log
ncu -i /tmp/adj.ncu-rep --page source --csv > /tmp/adj_sass.csv
ncu -i /tmp/adj.ncu-rep --page source --csv --print-source cuda > /tmp/adj_cuda.csv || true
python profiles/stall_by_line.py /tmp/adj_sass.csv 40 > gpurun_out/r02_adjoint_stalls_sass.txt
python profiles/stall_by_line.py /tmp/adj_cuda.csv 40 > gpurun_out/r02_adjoint_stalls_cuda.txt || true
head -c 1500 /tmp/adj_sass.csv > gpurun_out/adj_sass_head.txt
