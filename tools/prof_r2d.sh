set -x
python bench.py --no-sub-records --no-cpu-baseline --steps 1 --warmup 1 > gpurun_out/r2d_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 1700 -c 1700 --csv --log-file gpurun_out/r2d_launches_m32k.csv python bench.py --no-sub-records --no-cpu-baseline --steps 1 --warmup 1 > gpurun_out/r2d_ncu_bench.log 2>&1
python tools/one_eval.py 32768 1 > gpurun_out/r2d_plain_one.log 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'gemm_kernel.*GemmCfg<64, 128.*GeoSyrk' -s 0 -c 4 -o gpurun_out/r2d_syrk python tools/one_eval.py 32768 1 > gpurun_out/r2d_ncu_one.log 2>&1
ls -la gpurun_out/
