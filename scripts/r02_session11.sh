#!/bin/bash
# bench.py after the NUMA-placement change: default line, twice (e2e variance), fields of interest.
R=${1:-r02k}
mkdir -p gpurun_out
for i in 1 2; do
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${R}_bench$i.json 2> gpurun_out/${R}_bench$i.err; echo "bench rc=$?"; tail -2 gpurun_out/${R}_bench$i.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${R}_bench$i.json").read().splitlines()[-1])
print("value %.4e e2e %s ms %.3f" % (d["value"], d["e2e"], d["ms_per_step"]))
print("cpu", d["cpu_baseline"])
PY
done
lscpu | grep -i "numa\|socket\|model name" | head; cat /sys/bus/pci/devices/*/numa_node 2>/dev/null | sort | uniq -c | head
