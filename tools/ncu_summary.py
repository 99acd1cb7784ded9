"""Summarise an `ncu --csv --log-file` metrics pass (long format: one row per launch x metric).
usage: python tools/ncu_summary.py metrics.csv out_prefix [last_n]
writes <out_prefix>_summary.csv (one row per launch) and <out_prefix>_by_kernel.csv (grouped, with time share).
`last_n` keeps only the last n launches (drops warm-up iterations)."""
import csv, re, sys
from collections import OrderedDict, defaultdict

src, prefix = sys.argv[1], sys.argv[2]
last_n = int(sys.argv[3]) if len(sys.argv) > 3 else 0
rows = [r for r in csv.reader(l for l in open(src, errors="replace") if l.startswith('"'))]
hdr = rows[0]
iI, iK, iM, iU, iV = (hdr.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value"))
launches = OrderedDict()
for r in rows[1:]:
    d = launches.setdefault(int(r[iI]), {"kernel": r[iK]})
    v = float(r[iV].replace(",", "")) if r[iV] not in ("", "n/a") else 0.0
    u = r[iU]
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
    d[r[iM]] = v * scale
items = list(launches.values())
if last_n:
    items = items[-last_n:]


def short(k):
    k = re.sub(r"\(.*", "", k).replace("void ", "").replace("adp::", "")
    return k[:48]


cols = ("us", "dram_MB", "l2_MB", "tensor_active_pct", "lts_pct")
def rec(d):
    return (d.get("gpu__time_duration.sum", 0.0), d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0),
            d.get("lts__t_bytes.sum", 0.0), d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0),
            d.get("lts__throughput.avg.pct_of_peak_sustained_elapsed", 0.0))


with open(prefix + "_summary.csv", "w") as f:
    f.write("id,kernel," + ",".join(cols) + "\n")
    for i, d in enumerate(items):
        f.write(f"{i},{short(d['kernel'])}," + ",".join(f"{x:.1f}" for x in rec(d)) + "\n")
grp = defaultdict(lambda: [0, 0.0, 0.0, 0.0, 0.0])
for d in items:
    us, dram, l2, ta, _ = rec(d)
    g = grp[short(d["kernel"])]
    g[0] += 1; g[1] += us; g[2] += dram; g[3] += l2; g[4] += ta * us
tot = sum(g[1] for g in grp.values()) or 1.0
with open(prefix + "_by_kernel.csv", "w") as f:
    f.write("kernel,launches,total_us,share,dram_MB,l2_MB,tensor_active_pct_time_weighted\n")
    for k, g in sorted(grp.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k},{g[0]},{g[1]:.1f},{g[1] / tot:.3f},{g[2]:.1f},{g[3]:.1f},{g[4] / g[1] if g[1] else 0:.1f}\n")
    f.write(f"TOTAL,{sum(g[0] for g in grp.values())},{tot:.1f},1.000,{sum(g[2] for g in grp.values()):.1f},"
            f"{sum(g[3] for g in grp.values()):.1f},{sum(g[4] for g in grp.values()) / tot:.1f}\n")
print(open(prefix + "_by_kernel.csv").read())
