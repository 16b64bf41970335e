#!/bin/bash
# One GPU session: tests, smoke, bench (both arms), ncu launch list + full capture of the top kernels.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 600 gpurun_out/bench_ref.json
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
python bench.py --steps 2 --warmup 3 > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu1.log 2>&1
echo "ncu launches rc=$?"
python bench.py --steps 2 --warmup 3 > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"spread_kernel|gather_kernel" -s 6 -c 2 -o gpurun_out/prof_window -f python bench.py --steps 2 --warmup 3 > gpurun_out/ncu2.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/ncu2.log
