#!/bin/bash
# Round-2 GPU session 24: pruned transforms for 2D grids (c3) with the two-rows-per-transform X pass: on / off.
R=${1:-r02y}
mkdir -p gpurun_out
: > gpurun_out/${R}_ab.txt
run() {
  v=$(env $2 timeout 120 python bench.py --workload $1 --steps 10 --warmup 3 --no-extras 2>>gpurun_out/${R}_ab.err |
      python -c "import json,sys; d=json.loads(sys.stdin.read().replace('NaN','null')); print('%.4e %.3f ms %s' % (d['value'], d['ms_per_step'], json.dumps(d['stage_ms_per_step'])))")
  echo "$1 $2 $v" | tee -a gpurun_out/${R}_ab.txt
}
for rep in 1 2; do
for E in X=1 NFFTB200_PRUNED_2D=1; do run c3 $E; done
done
tail -5 gpurun_out/${R}_ab.err
