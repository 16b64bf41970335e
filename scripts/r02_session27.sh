#!/bin/bash
# Round-2 GPU session 27: resident CTAs per SM of the radix scatter kernel (NFFT_SORT_MINB 4 / 5 / 6 / 8): sort stage of c4.
R=${1:-r03d}
mkdir -p gpurun_out
: > gpurun_out/${R}_ab.txt
for rep in 1 2; do
for f in gpurun_variants/lib_*.so; do
  v=$(NFFTB200_LIB=$PWD/$f timeout 120 python bench.py --workload c4 --steps 10 --warmup 3 --no-extras 2>>gpurun_out/${R}_ab.err |
      python -c "import json,sys; d=json.loads(sys.stdin.read().replace('NaN','null')); print('%.4e %.3f ms %s' % (d['value'], d['ms_per_step'], json.dumps(d['stage_ms_per_step'])))")
  echo "c4 $f $v" | tee -a gpurun_out/${R}_ab.txt
done
done
tail -3 gpurun_out/${R}_ab.err
