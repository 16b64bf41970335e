#!/bin/bash
# Multi-GPU session: gpurun --gpus N -- bash scripts/r02_multi.sh <tag> <N>
R=${1:-r02m}; N=${2:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR scripts/dist_check.py 2>gpurun_out/${R}_n${N}_dist_check.err | tee gpurun_out/${R}_n${N}_dist_check.txt; tail -3 gpurun_out/${R}_n${N}_dist_check.err
timeout 900 $TR bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/${R}_n${N}_bench.json 2> gpurun_out/${R}_n${N}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${R}_n${N}_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/${R}_n${N}_bench.json"))
print("N=%d value %.4e e2e %.4e ms %.3f" % (d["n_gpus"], d["value"], d["e2e"]["value"], d["ms_per_step"]))
print("strong_c4", json.dumps(d.get("strong_c4")))
print("c5", json.dumps(d.get("c5_point_sharded")))
PY
for L2N in 26; do C5_LOG2N=$L2N timeout 300 $TR scripts/time_c5.py 2>>gpurun_out/${R}_n${N}_bench.err | tee -a gpurun_out/${R}_n${N}_c5.txt; done
