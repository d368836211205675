#!/bin/bash
# run the headline kernel timing for each tuning variant built by tools/build_variant.sh
cd "$(dirname "$0")/.."
for v in "$@"; do
  lib=hand-pose-estimation_b200/build/variants/libhpe_b200_$v.so
  HPE_B200_LIB=$PWD/$lib python bench.py --steps 300 --warmup 5 --no-extras 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('$v', 'evals/s %.4g  ms %.4f  frac %.3f  e2e %.4g' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value']))"
done
