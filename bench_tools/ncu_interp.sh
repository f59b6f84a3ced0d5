#!/bin/bash
# ncu --set full of the marching / plain interpolation + shift kernels at C2 size (bench_tools/quick_interp.py)
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'interp3|shift_window' -c 14 \
  -o gpurun_out/r02_interp_final -f python bench_tools/quick_interp.py --once > gpurun_out/r02_interp_ncu.log 2>&1
tail -2 gpurun_out/r02_interp_ncu.log
ls -la gpurun_out/*.ncu-rep
