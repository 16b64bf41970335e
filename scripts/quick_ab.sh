#!/bin/bash
# Short GPU session: A/B of gpurun_variants/lib_*.so on one workload, then the 3D parity tests with the fastest.
R=${1:-r01f}; WL=${2:-c4}
mkdir -p gpurun_out; : > gpurun_out/${R}_ab.txt
for f in gpurun_variants/lib_*.so; do
  v=$(NFFTB200_LIB=$PWD/$f timeout 60 python bench.py --workload $WL --steps 8 --warmup 3 --no-extras 2>>gpurun_out/${R}_ab.err |
      python -c "import json,sys; d=json.loads(sys.stdin.read().replace('NaN','null')); print('%.4e %s' % (d['value'], json.dumps(d['stage_ms_per_step'])))")
  echo "$f $v" | tee -a gpurun_out/${R}_ab.txt
done
BEST=$(sort -k2 -g -r gpurun_out/${R}_ab.txt | head -1 | cut -d' ' -f1)
echo "best variant: $BEST" | tee -a gpurun_out/${R}_ab.txt
NFFTB200_LIB=$PWD/$BEST timeout 70 python -m pytest tests/test_parity_gpu.py -m gpu -x -q > gpurun_out/${R}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${R}_ab.txt; tail -3 gpurun_out/${R}_pytest_gpu.log
