#!/bin/bash
# One GPU session that (1) A/B-times the library variants in gpurun_variants/ on a workload (default c4), (2) runs the GPU test
# suite, the bench and the ncu passes with the fastest one.  Usage (under gpurun): scripts/final_round.sh [round tag]
R=${1:-r01b}
WL=${2:-c4}   # workload the variants are compared on
mkdir -p gpurun_out
: > gpurun_out/${R}_ab.txt
for f in gpurun_variants/lib_*.so; do
  for rep in 1; do
    v=$(NFFTB200_LIB=$PWD/$f python bench.py --workload $WL --steps 8 --warmup 3 --no-extras 2>>gpurun_out/${R}_ab.err |
        python -c "import json,sys; d=json.loads(sys.stdin.read().replace('NaN','null')); print('%.4e %s' % (d['value'], json.dumps(d['stage_ms_per_step'])))")
    echo "$f $v" | tee -a gpurun_out/${R}_ab.txt
  done
done
BEST=$(python - <<PY
import collections
best = collections.defaultdict(float)
for line in open("gpurun_out/${R}_ab.txt"):
    p = line.split()
    if len(p) >= 2:
        try: best[p[0]] = max(best[p[0]], float(p[1]))
        except ValueError: pass
base = [k for k in best if "base" in k]
top = max(best, key=best.get)
if base and best[top] < 1.005 * best[base[0]]:
    top = base[0]
print(top)
PY
)
echo "best variant: $BEST" | tee -a gpurun_out/${R}_ab.txt
export NFFTB200_LIB=$PWD/$BEST
timeout 480 python -m pytest tests -m gpu -x -q > gpurun_out/${R}_pytest_gpu.log 2>&1; RC=$?; echo "pytest rc=$RC" | tee -a gpurun_out/${R}_ab.txt; tail -4 gpurun_out/${R}_pytest_gpu.log
BASE=$(ls gpurun_variants/lib_*base*.so 2>/dev/null | head -1)
if [ $RC -ne 0 ] && [ -n "$BASE" ] && [ "$BEST" != "$BASE" ]; then
  # the fastest variant is wrong: everything below describes the base build instead
  echo "falling back to $BASE" | tee -a gpurun_out/${R}_ab.txt
  export NFFTB200_LIB=$PWD/$BASE
  timeout 480 python -m pytest tests -m gpu -x -q > gpurun_out/${R}_pytest_gpu_base.log 2>&1; echo "pytest (base) rc=$?" | tee -a gpurun_out/${R}_ab.txt; tail -4 gpurun_out/${R}_pytest_gpu_base.log
fi
python bench.py --steps 10 --warmup 3 > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err; cat gpurun_out/${R}_bench.json
CMD="python bench.py --steps 2 --warmup 3 --no-extras"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches.csv $CMD > gpurun_out/${R}_ncu_launches.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"spread|gather" -s 6 -c 2 -o gpurun_out/${R}_window -f $CMD > gpurun_out/${R}_ncu_window.log 2>&1
echo "full capture rc=$?"
[ -n "$QUICK" ] || python bench.py --workload c2 --steps 50 --warmup 5 --no-extras --cuda-graph > gpurun_out/${R}_bench_c2_graph.json 2>> gpurun_out/${R}_bench.err; cat gpurun_out/${R}_bench_c2_graph.json
[ -n "$QUICK" ] || python bench.py --workload c3 --steps 8 --warmup 3 --no-extras > gpurun_out/${R}_bench_c3.json 2>> gpurun_out/${R}_bench.err; cut -c1-300 gpurun_out/${R}_bench_c3.json
python bench.py --workload c4_clustered --steps 8 --warmup 3 --no-extras > gpurun_out/${R}_bench_c4_clustered.json 2>> gpurun_out/${R}_bench.err; cut -c1-300 gpurun_out/${R}_bench_c4_clustered.json
[ -n "$QUICK" ] || python bench.py --impl reference --steps 2 --warmup 1 --ref-points 1048576 > gpurun_out/${R}_bench_ref.json 2> gpurun_out/${R}_bench_ref.err; cut -c1-600 gpurun_out/${R}_bench_ref.json
C5_LOG2N=23 python scripts/time_c5.py 2>>gpurun_out/${R}_bench.err | tee gpurun_out/${R}_c5.txt
[ -n "$QUICK" ] || C5_LOG2N=26 python scripts/time_c5.py 2>>gpurun_out/${R}_bench.err | tee -a gpurun_out/${R}_c5.txt
