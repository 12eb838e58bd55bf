# developer tool: the distributed default workload on N GPUs under the ownership widths (run under gpurun --gpus N)
N=${1:-2}
python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2d_pytest_multi_${N}gpu.log 2>&1; tail -2 gpurun_out/r2d_pytest_multi_${N}gpu.log
for ow in 2 1 4; do
  GPB_DIST_OW=$ow python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 2 --no-sub-records --no-cpu-baseline > gpurun_out/r2d_dist_n${N}_ow${ow}.json 2> gpurun_out/r2d_dist_n${N}_ow${ow}.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2d_dist_n${N}_ow${ow}.json").read().strip().splitlines()[-1])
    print("N=${N} OW=${ow}", d["value"], d["ms_per_step"], d["stages_ms"], d["check"])
except Exception as e:
    print("N=${N} OW=${ow} failed", e)
PY
done
