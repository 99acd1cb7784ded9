"""Print selected metrics from an ncu report (run on the GPU box so that only the summary travels back)."""
import csv, subprocess, sys
rep = sys.argv[1]
keys = sys.argv[2:] or []
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = [i for i, h in enumerate(hdr) if any(k in h for k in keys)] if keys else range(len(hdr))
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")][:50])
    for i in want:
        print(f"   {hdr[i]} = {r[i]} {units[i]}")
