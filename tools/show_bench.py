"""Pretty-print the JSON line of a bench.py run (and its sub-records): python tools/show_bench.py gpurun_out/x.json"""
import json
import sys


def show(r, ind=""):
    print(ind, r.get("config", {}).get("workload"), "| value", round(r["value"], 4), r.get("unit"), "| ms/step",
          round(r["ms_per_step"], 3), "| e2e", round(r["e2e"]["value"], 4))
    print(ind, "  stages", {k: round(v, 3) for k, v in r["stages_ms"].items()})
    print(ind, "  chol", {k: round(v, 3) for k, v in r["cholesky"].items()}, "launches", r.get("gpu_launches"))
    for k in ("cpu_baseline", "specialised_kernels", "check"):
        if k in r:
            print(ind, " ", k, r[k])


d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
show(d)
print("roofline", {k: v for k, v in d["roofline"].items() if k in ("achieved", "peak", "frac", "kernel")})
print("clocks", d["clocks"])
for k, v in d.get("sub_records", {}).items():
    show(v, "   ")
