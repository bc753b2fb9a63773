"""Summaries of `ncu --csv` logs for profiles/ (no GPU needed).

  python tools/ncu_summarize.py traffic  gpurun_out/gemm_traffic.csv  profiles/rN_gemm_traffic.json
      -> DRAM bytes of the LAST forward's gemm_kernel launches (the driver script runs two forwards)
  python tools/ncu_summarize.py launches gpurun_out/launches.csv      profiles/rN_launch_shares.json
      -> per-kernel share of the summed gpu__time_duration over the captured launches
"""
import csv
import json
import re
import sys
from collections import defaultdict


def rows(path):
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    return list(csv.DictReader(lines))


def short(name):
    name = re.sub(r"<unnamed>::|void |\(anonymous namespace\)::", "", name)
    return re.sub(r"\(.*", "", name)


def traffic(src, dst):
    per = defaultdict(dict)
    for r in rows(src):
        per[int(r["ID"])][r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
        per[int(r["ID"])]["name"] = r["Kernel Name"]
    ids = sorted(per)
    half = ids[len(ids) // 2:]                   # second forward (warm L2 / instruction caches)
    rd = sum(per[i].get("dram__bytes_read.sum", 0.0) for i in half)
    wr = sum(per[i].get("dram__bytes_write.sum", 0.0) for i in half)
    tm = sum(per[i].get("gpu__time_duration.sum", 0.0) for i in half)
    out = {"dram_bytes_per_forward": rd + wr, "dram_read_bytes": rd, "dram_write_bytes": wr, "launches": len(half),
           "sum_kernel_time_us_under_ncu": tm / 1e3, "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,"
           "gpu__time_duration.sum -k regex:gemm_kernel python tools/ncu_targets.py --model (ResNet-50, batch 256, second forward)"}
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out))


def launches(src, dst):
    tot = defaultdict(float)
    cnt = defaultdict(int)
    for r in rows(src):
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        v = v / 1e3 if r["Metric Unit"] in ("ns", "nsecond") else v
        k = short(r["Kernel Name"])
        tot[k] += v
        cnt[k] += 1
    s = sum(tot.values())
    out = {"total_us": s, "kernels": [{"kernel": k, "launches": cnt[k], "us": round(tot[k], 2), "share": round(tot[k] / s, 4)}
                                      for k in sorted(tot, key=lambda k: -tot[k])]}
    json.dump(out, open(dst, "w"), indent=1)
    for e in out["kernels"][:12]:
        print("%6.1f%%  %5d x  %s" % (100 * e["share"], e["launches"], e["kernel"]))


if __name__ == "__main__":
    {"traffic": traffic, "launches": launches}[sys.argv[1]](sys.argv[2], sys.argv[3])
