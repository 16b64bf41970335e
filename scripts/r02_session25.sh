#!/bin/bash
# Round-2 GPU session 25: key kernel without per-point divisions / batch searches: stage times of c4, c3, c2, c4_clustered,
# then the GPU tests (the binning tests compare keys and permutation bit for bit).
R=${1:-r03a}
mkdir -p gpurun_out
: > gpurun_out/${R}_ab.txt
run() {
  v=$(env $2 timeout 120 python bench.py --workload $1 --steps 10 --warmup 3 --no-extras 2>>gpurun_out/${R}_ab.err |
      python -c "import json,sys; d=json.loads(sys.stdin.read().replace('NaN','null')); print('%.4e %.3f ms %s' % (d['value'], d['ms_per_step'], json.dumps(d['stage_ms_per_step'])))")
  echo "$1 $2 $v" | tee -a gpurun_out/${R}_ab.txt
}
for WL in c4 c4_clustered c3 c2; do run $WL X=1; done
C5_LOG2N=26 timeout 120 python scripts/time_c5.py 2>>gpurun_out/${R}_ab.err | tee -a gpurun_out/${R}_ab.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee -a gpurun_out/${R}_ab.txt
tail -5 gpurun_out/${R}_ab.err
