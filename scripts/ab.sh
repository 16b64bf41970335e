#!/bin/bash
# A/B timing of library variants built into gpurun_variants/lib_*.so; usage: ab.sh [workload ...]
WL=${@:-c4}
for w in $WL; do
for f in gpurun_variants/lib_*.so; do
  n=$(basename $f .so)
  r=$(NFFTB200_LIB=$PWD/$f python bench.py --workload $w --steps 5 --warmup 3 --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.3e'%d['value'], d['stage_ms_per_step'])")
  echo "$w $n $r"
done
done
