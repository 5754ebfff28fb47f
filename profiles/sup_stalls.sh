set -e
ncu --set full --clock-control none --import-source on -k regex:cude_sup_kernel -s ${SKIP:-0} -c 1 -o /tmp/sup python profiles/ncu_targets.py sup_grad > /tmp/ncu.log 2>&1 || tail -5 /tmp/ncu.log
ncu -i /tmp/sup.ncu-rep --page source --csv > /tmp/sup_sass.csv
ncu -i /tmp/sup.ncu-rep --page raw --csv > /tmp/sup_raw.csv
python profiles/summarize_ncu.py /tmp/sup_raw.csv > gpurun_out/r02b_sup_grad_ncu_summary.txt
python profiles/stall_by_line.py /tmp/sup_sass.csv 45 > gpurun_out/r02b_sup_grad_stalls_sass.txt
