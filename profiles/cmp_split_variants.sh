#!/bin/bash
# ncu launch lists (gpu__time_duration) of one split-pipeline call for the product build and every tuning variant in
# conditional_ude_b200/csrc/variants/: profiles/cmp_split_variants.sh [individuals] [starts]
N=${1:-250000}; S=${2:-16}
mkdir -p gpurun_out
for lib in conditional_ude_b200/csrc/libcude_b200.so conditional_ude_b200/csrc/variants/*.so; do
  name=$(basename $lib .so)
  CUDE_B200_LIB=$PWD/$lib ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/var_$name.csv \
      python profiles/split_once.py $N $S 2 > gpurun_out/var_$name.log 2>&1
  echo "== $name"; python profiles/launch_table.py gpurun_out/var_$name.csv | head -4
done
