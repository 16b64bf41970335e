#!/bin/bash
# Round-2 GPU session 9: row skipping in the point slots: sanity, 3D parity tests, A/B on c4 / c4_clustered.
R=${1:-r02i}
mkdir -p gpurun_out
timeout 120 python scripts/tma_sanity.py > gpurun_out/${R}_tma_sanity.log 2>&1; echo "sanity rc=$?"; tail -3 gpurun_out/${R}_tma_sanity.log
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_parity_reference_gpu.py -m gpu -x -q > gpurun_out/${R}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${R}_pytest_gpu.log
: > gpurun_out/${R}_ab.txt
for WL in c4 c4_clustered; do
for f in gpurun_variants/lib_*.so; do
  for rep in 1 2; do
  v=$(NFFTB200_LIB=$PWD/$f timeout 120 python bench.py --workload $WL --steps 8 --warmup 3 --no-extras 2>>gpurun_out/${R}_ab.err |
      python -c "import json,sys; d=json.loads(sys.stdin.read().replace('NaN','null')); print('%.4e %s' % (d['value'], json.dumps(d['stage_ms_per_step'])))")
  echo "$WL $f $v" | tee -a gpurun_out/${R}_ab.txt
  done
done
done
