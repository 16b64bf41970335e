#!/bin/bash
# Final evidence, call D: one ncu --set full capture of the two FFT row kernels (the HBM-bound kernels of the path),
# after the same command exited 0 without ncu.
R=${1:-r03c}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-extras"
$CMD > gpurun_out/${R}_plain3.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${R}_plain3.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:"rows_r2c|rows_c2r" -s 6 -c 2 -o gpurun_out/${R}_fftrows -f $CMD > gpurun_out/${R}_ncu_fftrows.log 2>&1
echo "full capture rc=$?"; ls -la gpurun_out/${R}_fftrows.ncu-rep
