"""Where a kernel's executed instructions and stall samples go, from an `ncu --set full --import-source on` report.

    ncu -i gpurun_out/prof_layer0.ncu-rep --page source --csv --kernel-name regex:preprocess_tma > src.csv
    python tools/ncu_source_breakdown.py src.csv [window]

Prints the SASS stream in windows of `window` instructions (default 60): share of executed warp instructions, share of stall
samples, average active threads and the top opcodes of the window -- enough to tell polling loops, the main loop and the
epilogue apart without opening the GUI (this is how K1's 28 % of mbarrier polls were found).
"""
import collections
import csv
import sys


def main(path, window=60):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
    h = rows[hi]
    i_src, i_exec, i_samp, i_thr = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples"), h.index("Avg. Threads Executed")
    data = []
    for r in rows[hi + 1:]:
        if len(r) <= i_exec:
            continue
        try:
            data.append((r[i_src].strip(), int(r[i_exec]), int(r[i_samp]), float(r[i_thr])))
        except ValueError:
            continue
    tot = sum(d[1] for d in data) or 1
    ts = sum(d[2] for d in data) or 1
    print(f"{len(data)} SASS instructions, {tot} executed warp instructions, {ts} stall samples")
    ops_all = collections.Counter()
    for i in range(0, len(data), window):
        ch = data[i:i + window]
        e, s = sum(d[1] for d in ch), sum(d[2] for d in ch)
        ops = collections.Counter()
        for d in ch:
            tok = d[0].split()
            op = tok[1] if tok and tok[0].startswith("@") and len(tok) > 1 else (tok[0] if tok else "?")
            ops[op] += d[1]
            ops_all[op.split(".")[0]] += d[1]
        top = ", ".join(f"{k}:{v / tot * 100:.1f}" for k, v in ops.most_common(4))
        thr = sum(d[3] * d[1] for d in ch) / max(e, 1)
        print(f"{i:5d}  exec {e / tot * 100:5.1f}%  samples {s / ts * 100:5.1f}%  threads {thr:4.1f}  {top}")
    print("by opcode: " + ", ".join(f"{k} {v / tot * 100:.1f}%" for k, v in ops_all.most_common(12)))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 60)
