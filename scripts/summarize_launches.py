"""Summarises an `ncu --csv` launch list (gpu__time_duration + optional dram bytes) per kernel."""
import collections
import csv
import re
import sys


def main(path):
    path = str(path)
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    per = collections.defaultdict(lambda: collections.defaultdict(float))
    ids = collections.defaultdict(set)
    for r in rows:
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("peclr::", "")
        v = float(r["Metric Value"].replace(",", ""))
        m, unit = r["Metric Name"], r["Metric Unit"]
        if m == "gpu__time_duration.sum":
            v = v / 1e6 if unit in ("ns", "nsecond") else (v / 1e3 if unit in ("us", "usecond") else v)
            per[name]["ms"] += v
            ids[name].add(r["ID"])
        elif m.startswith("dram__bytes"):
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
            per[name]["bytes"] += v * scale
    tot = sum(v["ms"] for v in per.values())
    print("total %.3f ms over %d launches" % (tot, sum(len(s) for s in ids.values())))
    print("%-34s %5s %9s %6s %9s %8s" % ("kernel", "n", "ms", "share", "DRAM GB", "GB/s"))
    for k, v in sorted(per.items(), key=lambda kv: -kv[1]["ms"]):
        gb = v["bytes"] / 1e9
        print("%-34s %5d %9.3f %5.1f%% %9.2f %8.0f" % (k[:34], len(ids[k]), v["ms"], 100 * v["ms"] / tot, gb,
                                                      gb / (v["ms"] / 1e3) if v["ms"] else 0))


    if len(sys.argv) > 2:  # also write the per-launch DRAM traffic of the dominant kernel class for bench.py
        import json

        n = sum(len(ids[k]) for k in per if k.startswith("conv_gemm_kernel"))
        b = sum(v["bytes"] for k, v in per.items() if k.startswith("conv_gemm_kernel"))
        ms = sum(v["ms"] for k, v in per.items() if k.startswith("conv_gemm_kernel"))
        import os

        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        try:  # the build the capture was taken on (bench.py only quotes the traffic for that same build)
            digest = open(os.path.join(root, "peclr_b200", "csrc", ".build_stamp")).read().strip()[:16]
        except OSError:
            digest = None
        json.dump({"kernel": "conv_gemm_kernel", "launches": n, "dram_bytes_per_launch": b / max(n, 1),
                   "share_of_step_time": ms / tot, "source": path, "lib_digest": digest,
                   "workload": ["50", 128, 224]}, open(sys.argv[2], "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1])
