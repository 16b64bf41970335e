#!/bin/bash
# Round-2 GPU session 13: occupancy experiment (1 instead of 2 resident CTAs per SM through a shared-memory pad) + the
# dense-unit change on c5.
R=${1:-r02n}
mkdir -p gpurun_out
: > gpurun_out/${R}_ab.txt
for E in X=1 NFFTB200_SMEM_PAD=20000; do
  v=$(env $E timeout 120 python bench.py --workload c4 --steps 8 --warmup 3 --no-extras 2>>gpurun_out/${R}_ab.err |
      python -c "import json,sys; d=json.loads(sys.stdin.read().replace('NaN','null')); print('%.4e %s' % (d['value'], json.dumps(d['stage_ms_per_step'])))")
  echo "c4 $E $v" | tee -a gpurun_out/${R}_ab.txt
done
C5_LOG2N=23 timeout 120 python scripts/time_c5.py 2>>gpurun_out/${R}_ab.err | tee gpurun_out/${R}_c5.txt
C5_LOG2N=26 timeout 120 python scripts/time_c5.py 2>>gpurun_out/${R}_ab.err | tee -a gpurun_out/${R}_c5.txt
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -2
