#!/bin/bash
# Final evidence, call C: one ncu --set full capture of a spread and a gather launch, after the same command exited 0
# without ncu.
R=${1:-r02v}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-extras"
$CMD > gpurun_out/${R}_plain2.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${R}_plain2.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:"spread_reg|gather_reg" -s 6 -c 2 -o gpurun_out/${R}_window -f $CMD > gpurun_out/${R}_ncu_window.log 2>&1
echo "full capture rc=$?"; ls -la gpurun_out/${R}_window.ncu-rep
