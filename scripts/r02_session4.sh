#!/bin/bash
# Round-2 GPU session 4: (1) the standalone TMA probe, one test per process; (2) bisect of the c4 spread slowdown of
# session 3 with library variants (all with NFFTB200_NO_TMA=1 until the probe says which instruction faults).
R=${1:-r02d}
mkdir -p gpurun_out
: > gpurun_out/${R}_tma_probe.txt
for t in 0 1 2 3 4; do
  timeout 30 scripts/micro/tma_probe $t >> gpurun_out/${R}_tma_probe.txt 2>&1; echo "probe $t rc=$?" | tee -a gpurun_out/${R}_tma_probe.txt
done
cat gpurun_out/${R}_tma_probe.txt
export NFFTB200_NO_TMA=1
: > gpurun_out/${R}_ab.txt
for f in gpurun_variants/lib_*.so; do
  v=$(NFFTB200_LIB=$PWD/$f timeout 120 python bench.py --workload c4 --steps 8 --warmup 3 --no-extras 2>>gpurun_out/${R}_ab.err |
      python -c "import json,sys; d=json.loads(sys.stdin.read().replace('NaN','null')); print('%.4e %s' % (d['value'], json.dumps(d['stage_ms_per_step'])))")
  echo "c4 $f $v" | tee -a gpurun_out/${R}_ab.txt
done
NFFTB200_LIB=$PWD/gpurun_variants/lib_a_base.so timeout 120 python bench.py --workload c2 --steps 50 --warmup 5 --no-extras --cuda-graph 2>>gpurun_out/${R}_ab.err | cut -c1-200
NFFTB200_LIB=$PWD/gpurun_variants/lib_a_base.so timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -3
tail -5 gpurun_out/${R}_ab.err
