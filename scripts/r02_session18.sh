#!/bin/bash
# Round-2 GPU session 18: mixed-density mode behind the CLUSTERED hint: new tests, A/B of the heavy-tile threshold on the
# clustered workload, cost of the mode on the uniform one, stage times of c3 / c2; then the full GPU tests, smoke and
# the default bench line.
R=${1:-r02s2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "mixed or binning or plan" 2>&1 | tail -15
: > gpurun_out/${R}_ab.txt
run() {  # workload, env assignment
  v=$(env $2 timeout 120 python bench.py --workload $1 --steps 8 --warmup 3 --no-extras 2>>gpurun_out/${R}_ab.err |
      python -c "import json,sys; d=json.loads(sys.stdin.read().replace('NaN','null')); print('%.4e %.3f ms %s' % (d['value'], d['ms_per_step'], json.dumps(d['stage_ms_per_step'])))")
  echo "$1 $2 $v" | tee -a gpurun_out/${R}_ab.txt
}
for E in NFFTB200_NO_MIXED=1 X=1 NFFTB200_DENSE_TILE_PTS=1024 NFFTB200_DENSE_TILE_PTS=512 NFFTB200_DENSE_TILE_PTS=3072; do run c4_clustered $E; done
for E in X=1 NFFTB200_MIXED=1; do run c4 $E; done
run c3 X=1
run c2 X=1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee -a gpurun_out/${R}_ab.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err; echo "bench rc=$?"; cut -c1-3000 gpurun_out/${R}_bench.json
tail -5 gpurun_out/${R}_ab.err
