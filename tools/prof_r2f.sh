(python tools/stage_sweep.py 1:8192 1:16384 1:32768 256:2048 1024:1024; GPB_SYRK_GROUPED=0 python tools/stage_sweep.py 1:8192 1:16384 1:32768 256:2048 1024:1024) > gpurun_out/grouped_sweep.jsonl 2>&1
python tools/one_eval.py 32768 1 > gpurun_out/r2f_plain_one.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 24 -c 1 -o gpurun_out/r2f_syrk python tools/one_eval.py 32768 1 > gpurun_out/r2f_ncu_one.log 2>&1
