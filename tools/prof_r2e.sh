# ncu --set full of the factorisation's trailing updates at n = 32768 (first outer step: k = 1024 bulk update)
python tools/one_eval.py 32768 1 > gpurun_out/r2e_plain_one.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 24 -c 10 -o gpurun_out/r2e_syrk python tools/one_eval.py 32768 1 > gpurun_out/r2e_ncu_one.log 2>&1
ls -la gpurun_out/ | grep r2e
