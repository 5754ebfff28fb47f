set -e
ncu --set full --clock-control none --import-source on -k regex:cude_warp_kernel -s 2 -c 1 -o /tmp/warp python profiles/warp_once.py 25 > /tmp/ncu.log 2>&1 || tail -5 /tmp/ncu.log
ncu -i /tmp/warp.ncu-rep --page source --csv > /tmp/warp_sass.csv
ncu -i /tmp/warp.ncu-rep --page raw --csv > /tmp/warp_raw.csv
python profiles/summarize_ncu.py /tmp/warp_raw.csv > gpurun_out/r02_warp_kernel_ncu_summary.txt
python profiles/stall_by_line.py /tmp/warp_sass.csv 30 > gpurun_out/r02_warp_kernel_stalls_sass.txt
