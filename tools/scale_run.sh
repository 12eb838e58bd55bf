#!/bin/bash
# Scaling sweep on one multi-GPU box: usage tools/scale_run.sh "<workloads>" "<gpu counts>" [extra bench args]
# Writes one JSON line per (workload, N) to gpurun_out/scale_<workload>_n<N>.json
WL=${1:-"c5s c3"}; NS=${2:-"1 2 4 8"}; shift 2
mkdir -p gpurun_out
port=29600
for w in $WL; do
  for n in $NS; do
    port=$((port+1))
    out=gpurun_out/scale_${w}_n${n}
    if [ "$n" = "1" ]; then
      timeout 900 python bench.py --gpus 1 --workload $w --no-cpu-baseline "$@" > $out.json 2> $out.err
    else
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
        bench.py --gpus $n --workload $w --no-cpu-baseline "$@" > $out.json 2> $out.err
    fi
    echo "$w N=$n rc=$? $(tail -c 300 $out.json | head -c 300)"
  done
done
