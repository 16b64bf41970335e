#!/bin/bash
# Round-2 GPU session 5: TMA (interior tiles) + dense 2x2x2 supercells + fine keys: probe, sanity, full tests, A/B, profiles.
R=${1:-r02e}
mkdir -p gpurun_out
: > gpurun_out/${R}_tma_probe.txt
for t in 5 6 7; do
  timeout 30 scripts/micro/tma_probe $t >> gpurun_out/${R}_tma_probe.txt 2>&1; echo "probe $t rc=$?" >> gpurun_out/${R}_tma_probe.txt
done
cat gpurun_out/${R}_tma_probe.txt
timeout 120 python scripts/tma_sanity.py > gpurun_out/${R}_tma_sanity.log 2>&1; RC=$?; echo "tma sanity rc=$RC"; tail -8 gpurun_out/${R}_tma_sanity.log
if [ $RC -ne 0 ]; then
  echo "TMA path failed: the rest of the session runs with NFFTB200_NO_TMA=1"
  export NFFTB200_NO_TMA=1
  timeout 120 python scripts/tma_sanity.py 2>&1 | tail -3
fi
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${R}_pytest_gpu.log 2>&1; PRC=$?; echo "pytest rc=$PRC"; tail -6 gpurun_out/${R}_pytest_gpu.log
: > gpurun_out/${R}_ab.txt
run() {  # label, workload, env...
  local label=$1 wl=$2; shift 2
  v=$(env "$@" timeout 120 python bench.py --workload $wl --steps 8 --warmup 3 --no-extras 2>>gpurun_out/${R}_ab.err |
      python -c "import json,sys; d=json.loads(sys.stdin.read().replace('NaN','null')); print('%.4e %s' % (d['value'], json.dumps(d['stage_ms_per_step'])))")
  echo "$wl $label $v" | tee -a gpurun_out/${R}_ab.txt
}
for WL in c4 c4_clustered; do
  run default $WL X=1
  run no_tma $WL NFFTB200_NO_TMA=1
done
for E in X=1 NFFTB200_NO_DENSE=1 NFFTB200_NO_TMA=1; do
  echo "c5 with $E" | tee -a gpurun_out/${R}_c5.txt
  env $E C5_LOG2N=23 timeout 120 python scripts/time_c5.py 2>>gpurun_out/${R}_ab.err | tee -a gpurun_out/${R}_c5.txt
  env $E C5_LOG2N=26 timeout 120 python scripts/time_c5.py 2>>gpurun_out/${R}_ab.err | tee -a gpurun_out/${R}_c5.txt
done
timeout 120 python bench.py --workload c2 --steps 50 --warmup 5 --no-extras --cuda-graph > gpurun_out/${R}_bench_c2_graph.json 2>>gpurun_out/${R}_ab.err
python -c "import json; d=json.load(open('gpurun_out/${R}_bench_c2_graph.json')); print('c2 eager ms', d['ms_per_step'], 'graph', d['cuda_graph'])"
if [ $PRC -eq 0 ]; then
  CMD="python bench.py --steps 2 --warmup 3 --no-extras"
  $CMD > gpurun_out/${R}_plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches.csv $CMD > gpurun_out/${R}_ncu_launches.log 2>&1
  echo "launch list rc=$?"
  ncu --set full --clock-control none --import-source on -k regex:"spread_reg|gather_reg" -s 6 -c 2 -o gpurun_out/${R}_window -f $CMD > gpurun_out/${R}_ncu_window.log 2>&1
  echo "full capture rc=$?"
fi
tail -5 gpurun_out/${R}_ab.err
