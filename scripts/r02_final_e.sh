#!/bin/bash
# Final evidence, call E: one ncu --set full capture of the binning and pack / unpack kernels (what bounds them),
# after the same command exited 0 without ncu.
R=${1:-r03c}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-extras"
$CMD > gpurun_out/${R}_plain4.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${R}_plain4.log; exit 1; }
ncu --set full --clock-control none -k regex:"key_tile|radix_scatter|radix_hist|unpack_kernel|pack_kernel|bin_search" -s 24 -c 7 -o gpurun_out/${R}_misc -f $CMD > gpurun_out/${R}_ncu_misc.log 2>&1
echo "full capture rc=$?"; ls -la gpurun_out/${R}_misc.ncu-rep
