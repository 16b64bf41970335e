#!/bin/bash
# Round-2 GPU session 19: row skipping in the 2D gather: A/B on c3, error against the fp64 oracle, GPU parity tests.
R=${1:-r02t}
mkdir -p gpurun_out
: > gpurun_out/${R}_ab.txt
for rep in 1 2; do
for f in gpurun_variants/lib_*.so; do
  v=$(NFFTB200_LIB=$PWD/$f timeout 120 python bench.py --workload c3 --steps 10 --warmup 3 --no-extras 2>>gpurun_out/${R}_ab.err |
      python -c "import json,sys; d=json.loads(sys.stdin.read().replace('NaN','null')); print('%.4e %.3f ms %s' % (d['value'], d['ms_per_step'], json.dumps(d['stage_ms_per_step'])))")
  echo "c3 $f $v" | tee -a gpurun_out/${R}_ab.txt
done
done
last=$(ls gpurun_variants/lib_*.so | tail -1)
NFFTB200_LIB=$PWD/$last timeout 300 python scripts/parity_probe.py 2>>gpurun_out/${R}_ab.err | tee -a gpurun_out/${R}_ab.txt
NFFTB200_LIB=$PWD/$last timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -3 | tee -a gpurun_out/${R}_ab.txt
tail -5 gpurun_out/${R}_ab.err
