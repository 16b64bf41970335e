#!/bin/bash
# GPU side of the profile artefacts (run under gpurun): plain run first, then the ncu passes of
# /opt/skills/guides/B200_PROFILING.md.  Outputs land in gpurun_out/; scripts/summarise_profiles.py
# (run on the CPU box) turns them into the committed files under profiles/.
set -x
R=${1:-r01}
CMD="python bench.py --steps 2 --warmup 3 --no-extras"
$CMD > gpurun_out/${R}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches.csv $CMD > gpurun_out/${R}_ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/${R}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"spread|gather" -s 6 -c 2 -o gpurun_out/${R}_window -f $CMD > gpurun_out/${R}_ncu_window.log 2>&1
echo "full capture rc=$?"
