"""Joins two time-only ncu launch lists of the same step launch by launch (scripts/ncu_ab_launches.sh)."""
import collections
import csv
import re
import sys


def load(path):
    rows = list(csv.DictReader(l for l in open(path) if not l.startswith("==")))
    out = collections.OrderedDict()
    for r in rows:
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("peclr::", "")
        e = out.setdefault(int(r["ID"]), {"name": name, "grid": r.get("Grid Size", "")})
        v = float(r["Metric Value"].replace(",", ""))
        if r["Metric Name"] == "gpu__time_duration.sum":
            e["us"] = v / 1e3 if r["Metric Unit"] in ("ns", "nsecond") else v
        else:
            e["smem"] = v * {"byte": 1, "Kbyte": 1e3}.get(r["Metric Unit"], 1)
    return list(out.values())


def main(pa, pb, la, lb):
    a, b = load(pa), load(pb)
    assert len(a) == len(b), (len(a), len(b))
    print("%-4s %-40s %-14s %9s %9s %8s   smem %s / %s" % ("#", "kernel", "grid", la[-12:], lb[-12:], "a-b us", la[-6:], lb[-6:]))
    per = collections.defaultdict(lambda: [0.0, 0.0, 0])
    for i, (x, y) in enumerate(zip(a, b)):
        assert x["name"] == y["name"], (i, x["name"], y["name"])
        per[x["name"]][0] += x["us"]; per[x["name"]][1] += y["us"]; per[x["name"]][2] += 1
        if x.get("smem") != y.get("smem") or abs(x["us"] - y["us"]) > 0.05 * max(x["us"], 5.0):
            print("%-4d %-40s %-14s %9.1f %9.1f %+8.1f   %6.0f / %6.0f" % (i, x["name"][:40], x["grid"], x["us"], y["us"],
                                                                     x["us"] - y["us"], x.get("smem", 0), y.get("smem", 0)))
    print()
    for k, (ua, ub, n) in sorted(per.items(), key=lambda kv: -kv[1][0]):
        print("%-40s %4d %10.1f %10.1f %+9.1f" % (k[:40], n, ua, ub, ua - ub))
    print("total %.1f %.1f" % (sum(v[0] for v in per.values()), sum(v[1] for v in per.values())))


if __name__ == "__main__":
    main(*sys.argv[1:5])
