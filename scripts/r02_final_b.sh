#!/bin/bash
# Final evidence, call B: the ncu launch list of one bench command, after the same command exited 0 without ncu.
R=${1:-r02v}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-extras"
$CMD > gpurun_out/${R}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${R}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches.csv $CMD > gpurun_out/${R}_ncu_launches.log 2>&1
echo "launch list rc=$?"
