"""Developer tool: condense ncu output into the tables kept under profiles/.
  ncu_summary.py full REPORT.ncu-rep OUT.csv   one row per captured launch of a `--set full` report
  ncu_summary.py launches LAUNCHES.csv OUT.txt per-kernel totals and shares of a `--metrics gpu__time_duration.sum` list
"""
import csv, collections, subprocess, sys

COLS = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum"]


def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(raw.splitlines()))
    h, units = r[0], r[1]
    idx = [h.index(c) for c in COLS]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(COLS)
        w.writerow([units[i] for i in idx])
        for row in r[2:]:
            w.writerow([row[i] for i in idx])


def launches(src, out, header=""):
    rows = [x for x in csv.reader(l for l in open(src) if not l.startswith("==")) if len(x) > 5]
    h = rows[0]
    kn, val, unit = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    tot = collections.Counter(); n = collections.Counter()
    for x in rows[1:]:
        v = float(x[val].replace(",", ""))
        v = v / 1000. if x[unit] in ("ns", "nsecond") else (v * 1000. if x[unit] in ("ms", "msecond") else v)
        tot[x[kn]] += v; n[x[kn]] += 1
    s = sum(tot.values())
    with open(out, "w") as f:
        if header:
            f.write(header + "\n")
        f.write("%-90s %5s %12s %7s\n" % ("kernel", "n", "total_us", "share"))
        for k, v in tot.most_common():
            f.write("%-90s %5d %12.1f %6.1f%%\n" % (k[:90], n[k], v, 100 * v / s))


if __name__ == "__main__":
    if sys.argv[1] == "full":
        full(sys.argv[2], sys.argv[3])
    else:
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
