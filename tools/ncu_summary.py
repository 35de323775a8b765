"""Compact per-launch table from an `ncu --page raw --csv` dump (the judged summary kept under profiles/).
    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > raw.csv ; python tools/ncu_summary.py raw.csv > profiles/<name>.md
"""
import csv
import sys

COLS = [("gpu__time_duration.sum", "time"), ("sm__cycles_elapsed.avg", "cycles"), ("dram__bytes_read.sum", "dram_rd"),
        ("dram__bytes_write.sum", "dram_wr"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%act"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"), ("smsp__inst_executed.sum", "warp_insts"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("lts__t_sector_hit_rate.pct", "l2hit%"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%")]


def main(path):
    rows = list(csv.reader(open(path)))
    h, units = rows[0], rows[1]
    ki = h.index("Kernel Name")
    print("| # | kernel | " + " | ".join(n for _, n in COLS) + " |")
    print("|---|---|" + "---|" * len(COLS))
    for n, r in enumerate(rows[2:]):
        cells = []
        for key, _ in COLS:
            if key in h:
                i = h.index(key)
                v = r[i]
                try:
                    f = float(v.replace(",", ""))
                    v = f"{f:.4g}" if abs(f) < 1e6 else f"{f:.4e}"
                except ValueError:
                    pass
                cells.append(f"{v} {units[i]}".strip())
            else:
                cells.append("-")
        name = r[ki].split("(")[0].replace("void ", "").replace("cre::", "")[:44]
        print(f"| {n} | `{name}` | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main(sys.argv[1])
