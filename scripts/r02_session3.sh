#!/bin/bash
# Round-2 GPU session 3: first contact of the TMA plane flush / load + fine sort keys + point-range units.
R=${1:-r02c}
mkdir -p gpurun_out
timeout 120 python scripts/tma_sanity.py > gpurun_out/${R}_tma_sanity.log 2>&1; RC=$?; echo "tma sanity rc=$RC"; tail -8 gpurun_out/${R}_tma_sanity.log
if [ $RC -ne 0 ]; then
  echo "TMA path failed its first contact: the rest of the session runs with NFFTB200_NO_TMA=1"
  export NFFTB200_NO_TMA=1
  timeout 120 python scripts/tma_sanity.py 2>&1 | tail -3
fi
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${R}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/${R}_pytest_gpu.log
: > gpurun_out/${R}_ab.txt
run() {  # label, workload, env...
  local label=$1 wl=$2; shift 2
  v=$(env "$@" timeout 120 python bench.py --workload $wl --steps 8 --warmup 3 --no-extras 2>>gpurun_out/${R}_ab.err |
      python -c "import json,sys; d=json.loads(sys.stdin.read().replace('NaN','null')); print('%.4e %s' % (d['value'], json.dumps(d['stage_ms_per_step'])))")
  echo "$wl $label $v" | tee -a gpurun_out/${R}_ab.txt
}
for WL in c4 c4_clustered; do
  run default $WL X=1
  run no_tma $WL NFFTB200_NO_TMA=1
  run no_fine $WL NFFTB200_NO_FINE_SORT=1
  run no_tma_no_fine $WL NFFTB200_NO_TMA=1 NFFTB200_NO_FINE_SORT=1
done
for E in X=1 NFFTB200_NO_TMA=1 NFFTB200_NO_FINE_SORT=1; do
  echo "c5 with $E"; env $E C5_LOG2N=23 timeout 120 python scripts/time_c5.py 2>>gpurun_out/${R}_ab.err | tee -a gpurun_out/${R}_c5.txt
  env $E C5_LOG2N=26 timeout 120 python scripts/time_c5.py 2>>gpurun_out/${R}_ab.err | tee -a gpurun_out/${R}_c5.txt
done
tail -5 gpurun_out/${R}_ab.err
