"""Per-kernel table of an .ncu-rep: launches, total / mean duration, DRAM bytes, issue utilisation, DRAM throughput.
usage: python scripts/ncu_kernel_table.py x.ncu-rep [particles]"""
import csv, io, subprocess, sys
from collections import OrderedDict
rep = sys.argv[1]
npart = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
def col(name):
    return hdr.index(name) if name in hdr else None
def val(r, name, scale_units=True):
    i = col(name)
    if i is None or r[i] in ("", "n/a"):
        return 0.0
    v = float(r[i].replace(",", ""))
    u = units[i]
    if scale_units:
        v *= {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(u, 1.0)
    return v
agg = OrderedDict()
for r in rows[2:]:
    name = r[col("Kernel Name")].split("(")[0].replace("void ", "").replace("nbk::", "")
    a = agg.setdefault(name, dict(n=0, ms=0.0, rd=0.0, wr=0.0, issue=0.0, inst=0.0, regs=0, dthr=0.0))
    ms = val(r, "gpu__time_duration.sum")
    a["n"] += 1; a["ms"] += ms; a["rd"] += val(r, "dram__bytes_read.sum"); a["wr"] += val(r, "dram__bytes_write.sum")
    a["issue"] += val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active", False) * ms
    a["dthr"] += val(r, "dram__throughput.avg.pct_of_peak_sustained_elapsed", False) * ms
    a["inst"] += val(r, "smsp__inst_executed.sum", False)
    a["regs"] = int(val(r, "launch__registers_per_thread", False))
tot = sum(a["ms"] for a in agg.values())
print("%-44s %5s %10s %7s %10s %10s %8s %8s %6s %5s" % ("kernel", "n", "total ms", "share", "DRAM rd MB", "DRAM wr MB", "GB/s", "issue %", "of HBM", "regs"))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
    gbs = (a["rd"] + a["wr"]) / (a["ms"] * 1e-3) / 1e9 if a["ms"] else 0
    print("%-44s %5d %10.3f %6.1f%% %10.1f %10.1f %8.0f %8.1f %6.1f %5d" % (k[:44], a["n"], a["ms"], 100 * a["ms"] / tot, a["rd"] / 1e6, a["wr"] / 1e6, gbs,
                                                                    a["issue"] / a["ms"] if a["ms"] else 0, 100 * gbs / 6550.1, a["regs"]))
    if npart:
        print("%-44s       %.1f B/particle DRAM, %.0f warp-instructions/particle" % ("", (a["rd"] + a["wr"]) / npart, a["inst"] / npart))
print("(of HBM: DRAM GB/s over the measured copy peak 6550.1 GB/s of MEASURED_PEAKS.json)")
print("total %.3f ms over %d launches" % (tot, sum(a["n"] for a in agg.values())))
