"""Top stall-sample SASS lines of an `ncu --page source --csv` export (scripts/ncu_pick.sh): python ncu_src_top.py file [n]"""
import csv
import sys


def main(fn, n=25):
    rows = list(csv.reader(open(fn)))
    head, data = rows[1], rows[2:]
    si, src, ie = head.index("# Samples"), head.index("Source"), head.index("Instructions Executed")
    tot = sum(int(r[si]) for r in data) or 1
    print(rows[0][1][:90], "| samples", tot, "| warp-instr", sum(int(r[ie]) for r in data))
    idx = sorted(range(len(data)), key=lambda i: -int(data[i][si]))[:n]
    for i in sorted(idx):
        print("%5d %6s %5.1f%% %9s  %s" % (i, data[i][si], 100 * int(data[i][si]) / tot, data[i][ie], data[i][src].strip()[:95]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
