#!/bin/bash
# Round-2 GPU session 2: the refactored library (point plans, cuFFT work areas, batch offsets): full GPU test suite, smoke,
# default bench line (with extra_workloads / c5), a short reference arm, then A/B of the lock / stagger variants.
R=${1:-r02b}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${R}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/${R}_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err; echo "bench rc=$?"; cat gpurun_out/${R}_bench.json; tail -3 gpurun_out/${R}_bench.err
timeout 300 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/${R}_bench_ref.json 2> gpurun_out/${R}_bench_ref.err; echo "ref rc=$?"; cut -c1-900 gpurun_out/${R}_bench_ref.json
: > gpurun_out/${R}_ab.txt
for WL in c4; do
for f in gpurun_variants/lib_*.so; do
  v=$(NFFTB200_LIB=$PWD/$f timeout 120 python bench.py --workload $WL --steps 8 --warmup 3 --no-extras 2>>gpurun_out/${R}_ab.err |
      python -c "import json,sys; d=json.loads(sys.stdin.read().replace('NaN','null')); print('%.4e %s' % (d['value'], json.dumps(d['stage_ms_per_step'])))")
  echo "$WL $f $v" | tee -a gpurun_out/${R}_ab.txt
done
done
tail -5 gpurun_out/${R}_ab.err
