#!/bin/bash
# ncu --set full of the 8-lane warp kernel (3990 trajectories) and of the two suppression gradient kernels (37 x 10 000)
set -e
ncu --set full --clock-control none -k regex:cude_warp_kernel -s 2 -c 1 -o /tmp/w8 python profiles/warp_once.py 70 > /tmp/n1.log 2>&1 || tail -3 /tmp/n1.log
ncu -i /tmp/w8.ncu-rep --page raw --csv > /tmp/w8.csv; python profiles/summarize_ncu.py /tmp/w8.csv > gpurun_out/r02_warp8_kernel_ncu_summary.txt
ncu --set full --clock-control none -k regex:cude_sup_kernel -c 2 -o /tmp/s2 python profiles/ncu_targets.py sup_grad > /tmp/n2.log 2>&1 || tail -3 /tmp/n2.log
ncu -i /tmp/s2.ncu-rep --page raw --csv > /tmp/s2.csv; python profiles/summarize_ncu.py /tmp/s2.csv > gpurun_out/r02c_sup_two_kernel_ncu_summary.txt
