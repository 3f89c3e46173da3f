"""Condenses an `ncu --set full` report to the metrics the profiles/ summaries quote:
    ncu -i report.ncu-rep --page raw --csv | python scripts/ncu_summary.py > profiles/<name>.txt"""
import csv
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__cluster_dim_x", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__cycles_active.avg", "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]


def main():
    rows = list(csv.reader(l for l in sys.stdin if not l.startswith("==")))
    head, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(head, r))
        u = dict(zip(head, units))
        print("----")
        print("  %-70s %s" % ("Kernel Name", d.get("Kernel Name", "").replace("void ", "").replace("peclr::", "")))
        print("  %-70s %s  grid %s" % ("Block / grid", d.get("Block Size"), d.get("Grid Size")))
        for k in KEEP:
            for name in head:
                if name == k or name.endswith("." + k):
                    if d.get(name, "") != "":
                        print("  %-70s %s %s" % (k, d[name], u.get(name, "")))
                    break


if __name__ == "__main__":
    main()
