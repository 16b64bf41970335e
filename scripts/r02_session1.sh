#!/bin/bash
# Round-2 GPU session 1: new golden fixtures from the compiled reference, the BASELINE-size parity tests against the
# reference, A/B of the queued kernel variants (c4 and c4_clustered), c5 timing.   gpurun -- bash scripts/r02_session1.sh
R=r02a
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 300 python tests/golden/make_golden.py --new-only > gpurun_out/${R}_golden.log 2>&1; echo "golden rc=$?"; tail -2 gpurun_out/${R}_golden.log
rm -f gpurun_out/parity_reference.jsonl
timeout 900 python -m pytest tests/test_parity_reference_gpu.py -m gpu -q -s > gpurun_out/${R}_parity_reference.log 2>&1; echo "parity-vs-reference rc=$?"; tail -5 gpurun_out/${R}_parity_reference.log
cat gpurun_out/parity_reference.jsonl
: > gpurun_out/${R}_ab.txt
for WL in c4 c4_clustered; do
for f in gpurun_variants/lib_*.so; do
  v=$(NFFTB200_LIB=$PWD/$f timeout 120 python bench.py --workload $WL --steps 8 --warmup 3 --no-extras 2>>gpurun_out/${R}_ab.err |
      python -c "import json,sys; d=json.loads(sys.stdin.read().replace('NaN','null')); print('%.4e %s' % (d['value'], json.dumps(d['stage_ms_per_step'])))")
  echo "$WL $f $v" | tee -a gpurun_out/${R}_ab.txt
done
done
C5_LOG2N=23 timeout 120 python scripts/time_c5.py 2>>gpurun_out/${R}_ab.err | tee gpurun_out/${R}_c5.txt
C5_LOG2N=26 timeout 120 python scripts/time_c5.py 2>>gpurun_out/${R}_ab.err | tee -a gpurun_out/${R}_c5.txt
tail -5 gpurun_out/${R}_ab.err
