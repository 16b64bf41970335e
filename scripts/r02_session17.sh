#!/bin/bash
# Round-2 GPU session 17: mixed-density mode (heavy tiles swept with 2 x 2 x 2 supercells, decided on the device):
# the new tests first, then c4 / c4_clustered with the mode off, on, and with other heavy-tile thresholds; full GPU tests.
R=${1:-r02r}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "mixed or binning or plan" 2>&1 | tail -15
: > gpurun_out/${R}_ab.txt
for WL in c4_clustered c4; do
for E in NFFTB200_NO_MIXED=1 X=1 NFFTB200_DENSE_TILE_PTS=4096 NFFTB200_DENSE_TILE_PTS=2048 NFFTB200_DENSE_TILE_PTS=6144; do
  v=$(env $E timeout 120 python bench.py --workload $WL --steps 8 --warmup 3 --no-extras 2>>gpurun_out/${R}_ab.err |
      python -c "import json,sys; d=json.loads(sys.stdin.read().replace('NaN','null')); print('%.4e %.3f ms %s' % (d['value'], d['ms_per_step'], json.dumps(d['stage_ms_per_step'])))")
  echo "$WL $E $v" | tee -a gpurun_out/${R}_ab.txt
done
done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee -a gpurun_out/${R}_ab.txt
tail -5 gpurun_out/${R}_ab.err
