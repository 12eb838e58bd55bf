# developer tool: C5 (n = 65536 likelihood) on N GPUs with replicated and with column storage
N=$1; NN=${2:-65536}
for st in replicated columns; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload c5 --size $NN --storage $st --steps 2 --warmup 1 --no-sub-records --no-cpu-baseline > gpurun_out/r2g_c5_n${N}_${st}.json 2> gpurun_out/r2g_c5_n${N}_${st}.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2g_c5_n${N}_${st}.json").read().strip().splitlines()[-1])
    print("N=${N} ${st}", round(d["value"],4), round(d["ms_per_step"],1), {k: round(x,1) for k,x in d["stages_ms"].items()}, d["check"]["nll0"], d["config"]["workspace_gib_per_rank"])
except Exception as e:
    print("N=${N} ${st} failed", e); print(open("gpurun_out/r2g_c5_n${N}_${st}.err").read()[-1500:])
PY
done
