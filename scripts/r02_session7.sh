#!/bin/bash
# Round-2 GPU session 7: validation of the consolidated build: GPU tests, smoke, default bench line, c5 stages, c4 A/B vs NO_TMA.
R=${1:-r02g}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${R}_pytest_gpu.log 2>&1; PRC=$?; echo "pytest rc=$PRC"; tail -4 gpurun_out/${R}_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/${R}_bench.json; tail -3 gpurun_out/${R}_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/${R}_bench.json"))
print("value %.4e e2e %.4e ms %.3f stages %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["stage_ms_per_step"]))
print("extra", {k: (round(v["ms_per_step"], 4), "%.3e" % v["value"]) for k, v in d.get("extra_workloads", {}).items()})
print("c2 graph", d.get("extra_workloads", {}).get("c2", {}).get("cuda_graph"))
print("c5", d.get("c5_point_sharded"))
PY
for E in X=1 NFFTB200_NO_TMA=1; do
  v=$(env $E timeout 120 python bench.py --workload c4 --steps 8 --warmup 3 --no-extras 2>>gpurun_out/${R}_ab.err |
      python -c "import json,sys; d=json.loads(sys.stdin.read().replace('NaN','null')); print('%.4e %s' % (d['value'], json.dumps(d['stage_ms_per_step'])))")
  echo "c4 $E $v" | tee -a gpurun_out/${R}_ab.txt
done
C5_LOG2N=23 timeout 120 python scripts/time_c5.py 2>>gpurun_out/${R}_ab.err | tee gpurun_out/${R}_c5.txt
C5_LOG2N=26 timeout 120 python scripts/time_c5.py 2>>gpurun_out/${R}_ab.err | tee -a gpurun_out/${R}_c5.txt
