#!/bin/bash
# Round-2 GPU session 12: units per batch (8 / 16 / 24) on dense and clustered data.
R=${1:-r02l}
mkdir -p gpurun_out
: > gpurun_out/${R}_ab.txt
for f in gpurun_variants/lib_*.so; do
  echo "c5 with $f" | tee -a gpurun_out/${R}_c5.txt
  NFFTB200_LIB=$PWD/$f C5_LOG2N=23 timeout 120 python scripts/time_c5.py 2>>gpurun_out/${R}_ab.err | tee -a gpurun_out/${R}_c5.txt
  NFFTB200_LIB=$PWD/$f C5_LOG2N=26 timeout 120 python scripts/time_c5.py 2>>gpurun_out/${R}_ab.err | tee -a gpurun_out/${R}_c5.txt
  for WL in c4_clustered c4; do
  v=$(NFFTB200_LIB=$PWD/$f timeout 120 python bench.py --workload $WL --steps 8 --warmup 3 --no-extras 2>>gpurun_out/${R}_ab.err |
      python -c "import json,sys; d=json.loads(sys.stdin.read().replace('NaN','null')); print('%.4e %s' % (d['value'], json.dumps(d['stage_ms_per_step'])))")
  echo "$WL $f $v" | tee -a gpurun_out/${R}_ab.txt
  done
done
