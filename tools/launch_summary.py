"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, mean, share."""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        full = row["Kernel Name"]
        name = re.sub(r"\(.*", "", full).replace("void ", "")
        geo = re.search(r"gpb::(Geo\w+)", full)
        if geo:   # gemm_kernel<GemmCfg<BM, BN, ..>, AKM, BKM, Geo>: keep the tile shape and the geometry
            dims = re.findall(r"\(int\)(\d+)", full)
            name = "gemm<%sx%s,%s>" % (dims[0], dims[1], geo.group(1)) if len(dims) >= 2 else "gemm<%s>" % geo.group(1)
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
        agg.setdefault(name, []).append(v)
    tot = sum(sum(v) for v in agg.values())
    print("%-44s %6s %10s %10s %10s %10s %7s" % ("kernel", "n", "sum_ms", "mean_us", "min_us", "max_us", "share"))
    for k, v in agg.items():
        print("%-44s %6d %10.3f %10.1f %10.1f %10.1f %6.1f%%" % (k[:44], len(v), sum(v) / 1e3, sum(v) / len(v), min(v),
                                                                 max(v), 100 * sum(v) / tot))
    print("total_ms %.3f (serialised, cold-cache launch times: compare shares, not absolutes)" % (tot / 1e3))


if __name__ == "__main__":
    main(sys.argv[1])
