# developer tool: distributed default workload on N GPUs, env variants given as "VAR=val,VAR=val" arguments
N=$1; shift
for v in "$@"; do
  tag=$(echo "$v" | tr -c 'A-Za-z0-9=\n' '_')
  env $(echo "$v" | tr ',' ' ') python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 2 --no-sub-records --no-cpu-baseline > gpurun_out/r2d_dist_n${N}_${tag}.json 2> gpurun_out/r2d_dist_n${N}_${tag}.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2d_dist_n${N}_${tag}.json").read().strip().splitlines()[-1])
    print("N=${N} ${v}", round(d["value"],4), round(d["ms_per_step"],1), {k: round(x,1) for k,x in d["stages_ms"].items()}, d["check"]["nll0"])
except Exception as e:
    print("N=${N} ${v} failed", e)
PY
done
