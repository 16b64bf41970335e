#!/bin/bash
# Round-2 GPU session 21: pruned real transforms (hand-written X pass + cuFFT C2C on the kept kx planes): the new
# test first, then c4 / c3 stage times with the path on and off, then the full GPU tests.
R=${1:-r02w}
mkdir -p gpurun_out
: > gpurun_out/${R}_ab.txt
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "pruned" 2>&1 | tail -8 | tee -a gpurun_out/${R}_ab.txt
grep -q " passed" gpurun_out/${R}_ab.txt && ! grep -q "failed\|error" gpurun_out/${R}_ab.txt || { echo "pruned test failed: stop"; exit 0; }
run() {
  v=$(env $2 timeout 120 python bench.py --workload $1 --steps 8 --warmup 3 --no-extras 2>>gpurun_out/${R}_ab.err |
      python -c "import json,sys; d=json.loads(sys.stdin.read().replace('NaN','null')); print('%.4e %.3f ms %s' % (d['value'], d['ms_per_step'], json.dumps(d['stage_ms_per_step'])))")
  echo "$1 $2 $v" | tee -a gpurun_out/${R}_ab.txt
}
for rep in 1 2; do
for WL in c4; do
for E in NFFTB200_NO_PRUNED_FFT=1 X=1; do run $WL $E; done
done
done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee -a gpurun_out/${R}_ab.txt
tail -5 gpurun_out/${R}_ab.err
