#!/bin/bash
# Round-2 GPU session 15: A/B of the library variants in gpurun_variants/ on c4 and c3 (stage times), their error against
# the fp64 oracle (scripts/parity_probe.py), and the GPU parity tests with the last variant.
R=${1:-r02p}
mkdir -p gpurun_out
: > gpurun_out/${R}_ab.txt
for WL in c4 c3; do
for f in gpurun_variants/lib_*.so; do
  v=$(NFFTB200_LIB=$PWD/$f timeout 120 python bench.py --workload $WL --steps 8 --warmup 3 --no-extras 2>>gpurun_out/${R}_ab.err |
      python -c "import json,sys; d=json.loads(sys.stdin.read().replace('NaN','null')); print('%.4e %s' % (d['value'], json.dumps(d['stage_ms_per_step'])))")
  echo "$WL $f $v" | tee -a gpurun_out/${R}_ab.txt
done
done
for f in gpurun_variants/lib_*.so; do
  echo "parity $f" | tee -a gpurun_out/${R}_ab.txt
  NFFTB200_LIB=$PWD/$f timeout 300 python scripts/parity_probe.py 2>>gpurun_out/${R}_ab.err | tee -a gpurun_out/${R}_ab.txt
done
last=$(ls gpurun_variants/lib_*.so | tail -1)
echo "tests with $last"
NFFTB200_LIB=$PWD/$last timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee -a gpurun_out/${R}_ab.txt
tail -5 gpurun_out/${R}_ab.err
