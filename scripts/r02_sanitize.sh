#!/bin/bash
# One compute-sanitizer tool per gpurun call, after the same command has exited 0 without it (B200_PROFILING.md).
R=${1:-r02s}
mkdir -p gpurun_out
timeout 120 python scripts/sanitize_case.py > gpurun_out/${R}_sanitize_plain.log 2>&1 &&
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 5 python scripts/sanitize_case.py > gpurun_out/${R}_memcheck.log 2>&1
echo "memcheck rc=$?"
tail -5 gpurun_out/${R}_sanitize_plain.log; tail -15 gpurun_out/${R}_memcheck.log
