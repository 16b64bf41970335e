#!/bin/bash
# Round-2 GPU session 8: final evidence of the consolidated build: tests, smoke, bench (both arms), ncu launch list + capture.
R=${1:-r02j}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${R}_pytest_gpu.log 2>&1; PRC=$?; echo "pytest rc=$PRC"; tail -4 gpurun_out/${R}_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${R}_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${R}_bench.json").read().splitlines()[-1])
print("value %.4e e2e %.4e ms %.3f stages %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["stage_ms_per_step"]))
print("roofline frac %.4f achieved %.1f" % (d["roofline"]["frac"], d["roofline"]["achieved"]))
print("extra", {k: (round(v["ms_per_step"], 4), "%.3e" % v["value"]) for k, v in d.get("extra_workloads", {}).items()})
print("c2 graph", d.get("extra_workloads", {}).get("c2", {}).get("cuda_graph"))
print("c5", d.get("c5_point_sharded"))
PY
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${R}_bench_ref.json 2> gpurun_out/${R}_bench_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/${R}_bench_ref.json
if [ $PRC -eq 0 ]; then
  CMD="python bench.py --steps 2 --warmup 3 --no-extras"
  $CMD > gpurun_out/${R}_plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches.csv $CMD > gpurun_out/${R}_ncu_launches.log 2>&1
  echo "launch list rc=$?"
  ncu --set full --clock-control none --import-source on -k regex:"spread_reg|gather_reg" -s 6 -c 2 -o gpurun_out/${R}_window -f $CMD > gpurun_out/${R}_ncu_window.log 2>&1
  echo "full capture rc=$?"
fi
