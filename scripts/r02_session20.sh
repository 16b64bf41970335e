#!/bin/bash
# Round-2 GPU session 20: rotating z block in the gather (no slide): A/B on c4 / c4_clustered / c5, parity probe, GPU tests.
R=${1:-r02u}
mkdir -p gpurun_out
: > gpurun_out/${R}_ab.txt
last=$(ls gpurun_variants/lib_*.so | tail -1)
NFFTB200_LIB=$PWD/$last timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -3 | tee -a gpurun_out/${R}_ab.txt
grep -q " passed" gpurun_out/${R}_ab.txt && ! grep -q "failed" gpurun_out/${R}_ab.txt || { echo "tests failed: no A/B"; exit 0; }
for rep in 1 2; do
for WL in c4 c4_clustered; do
for f in gpurun_variants/lib_*.so; do
  v=$(NFFTB200_LIB=$PWD/$f timeout 120 python bench.py --workload $WL --steps 8 --warmup 3 --no-extras 2>>gpurun_out/${R}_ab.err |
      python -c "import json,sys; d=json.loads(sys.stdin.read().replace('NaN','null')); print('%.4e %.3f ms %s' % (d['value'], d['ms_per_step'], json.dumps(d['stage_ms_per_step'])))")
  echo "$WL $f $v" | tee -a gpurun_out/${R}_ab.txt
done
done
done
for f in gpurun_variants/lib_*.so; do
  echo "c5 $f" | tee -a gpurun_out/${R}_ab.txt
  NFFTB200_LIB=$PWD/$f C5_LOG2N=26 timeout 120 python scripts/time_c5.py 2>>gpurun_out/${R}_ab.err | tee -a gpurun_out/${R}_ab.txt
done
last=$(ls gpurun_variants/lib_*.so | tail -1)
NFFTB200_LIB=$PWD/$last timeout 300 python scripts/parity_probe.py 2>>gpurun_out/${R}_ab.err | tee -a gpurun_out/${R}_ab.txt
NFFTB200_LIB=$PWD/$last timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee -a gpurun_out/${R}_ab.txt
tail -5 gpurun_out/${R}_ab.err
