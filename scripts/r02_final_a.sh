#!/bin/bash
# Final evidence, call A: GPU tests, smoke, the default bench line (both arms).  No profiler in this call.
R=${1:-r02v}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${R}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${R}_pytest_gpu.log
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
print("cpu_baseline", d.get("cpu_baseline"))
PY
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${R}_bench_ref.json 2> gpurun_out/${R}_bench_ref.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/${R}_bench_ref.json
C5_LOG2N=26 timeout 120 python scripts/time_c5.py 2>>gpurun_out/${R}_bench.err | tee gpurun_out/${R}_c5.txt
