#!/bin/bash
# Round-2 GPU session 6: multi-batch work items; A/B of batches 1/4/8 and the old strides; c5 stages.
R=${1:-r02f}
mkdir -p gpurun_out
timeout 120 python scripts/tma_sanity.py > gpurun_out/${R}_tma_sanity.log 2>&1; RC=$?; echo "tma sanity rc=$RC"; tail -8 gpurun_out/${R}_tma_sanity.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${R}_pytest_gpu.log 2>&1; PRC=$?; echo "pytest rc=$PRC"; tail -6 gpurun_out/${R}_pytest_gpu.log
: > gpurun_out/${R}_ab.txt
for WL in c4 c4_clustered; do
for f in gpurun_variants/lib_*.so; do
  v=$(NFFTB200_LIB=$PWD/$f timeout 120 python bench.py --workload $WL --steps 8 --warmup 3 --no-extras 2>>gpurun_out/${R}_ab.err |
      python -c "import json,sys; d=json.loads(sys.stdin.read().replace('NaN','null')); print('%.4e %s' % (d['value'], json.dumps(d['stage_ms_per_step'])))")
  echo "$WL $f $v" | tee -a gpurun_out/${R}_ab.txt
done
done
for f in gpurun_variants/lib_a_base.so gpurun_variants/lib_b_batch1.so gpurun_variants/lib_d_batch8.so; do
  echo "c5 with $f" | tee -a gpurun_out/${R}_c5.txt
  NFFTB200_LIB=$PWD/$f C5_LOG2N=23 timeout 120 python scripts/time_c5.py 2>>gpurun_out/${R}_ab.err | tee -a gpurun_out/${R}_c5.txt
  NFFTB200_LIB=$PWD/$f C5_LOG2N=26 timeout 120 python scripts/time_c5.py 2>>gpurun_out/${R}_ab.err | tee -a gpurun_out/${R}_c5.txt
done
tail -5 gpurun_out/${R}_ab.err
