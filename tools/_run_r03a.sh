mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:harmonic_dft -s 1 -c 1 -o gpurun_out/r03g_harmonic_dft32 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r03g_ncu.log 2>&1
tail -3 gpurun_out/r03g_ncu.log
