"""One line per kernel of a multi-kernel .ncu-rep: duration, DRAM bytes read / written, DRAM and L2 throughput.
usage: python tools/ncu_table.py <report.ncu-rep> "<header comment>" > profiles/<name>.txt"""
import csv
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]


def col(name):
    return hdr.index(name) if name in hdr else None


def num(row, name, scale_by_unit=True):
    i = col(name)
    if i is None or row[i] in ("", "n/a"):
        return float("nan")
    v = float(row[i].replace(",", ""))
    u = units[i].lower()
    if scale_by_unit:
        for pre, f in (("gbyte", 1e9), ("mbyte", 1e6), ("kbyte", 1e3), ("msecond", 1e-3), ("usecond", 1e-6),
                       ("nsecond", 1e-9), ("ms", 1e-3), ("us", 1e-6), ("ns", 1e-9)):
            if u.startswith(pre):
                return v * f
    return v


print("# " + sys.argv[2])
print("# %-44s %10s %10s %10s %9s %7s %7s %5s" % ("kernel", "time us", "rd MB", "wr MB", "DRAM GB/s", "dram %", "l2 %",
                                                 "regs"))
for r in rows[2:]:
    name = r[col("Kernel Name")][:44]
    t = num(r, "gpu__time_duration.sum")
    rd, wr = num(r, "dram__bytes_read.sum"), num(r, "dram__bytes_write.sum")
    print("  %-44s %10.1f %10.1f %10.1f %9.0f %7.1f %7.1f %5.0f" % (
        name, t * 1e6, rd / 1e6, wr / 1e6, (rd + wr) / t / 1e9,
        num(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", False),
        num(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed", False),
        num(r, "launch__registers_per_thread", False)))
