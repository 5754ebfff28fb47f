#!/bin/bash
# bash profiles/cmp_variants.sh main v7 b64mb6:64 ... : bench.py (1M x 64, FP64) with libcude_b200.so or
# csrc/variants/<name>.so, optional :<block size>
for spec in "$@"; do
  v=${spec%%:*}; b=0; [[ "$spec" == *:* ]] && b=${spec#*:}
  echo -n "$spec: "
  if [ "$v" = "main" ]; then L=/root/repo/conditional_ude_b200/csrc/libcude_b200.so; else L=/root/repo/conditional_ude_b200/csrc/variants/$v.so; fi
  CUDE_B200_LIB=$L python bench.py --steps 3 --warmup 3 --no-cpu-baseline --block $b 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.4g evals/s  kernel_ms %.2f  frac %.3f lossonly %.4g e2e %.4g'%(d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['config']['loss_only_evals_per_s'], d['e2e']['value']))"
done
