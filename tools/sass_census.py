"""Per-kernel census of the Blackwell-specific SASS mnemonics in libadapose_b200.so (cuobjdump -sass), written as CSV.
usage: python tools/sass_census.py [out.csv]"""
import os, re, subprocess, sys
from collections import Counter, OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "rgbmanip_b200", "libadapose_b200.so")
COLS = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCOMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "UTCCP", "SYNCS", "MUFU.EX2", "HMMA",
        "IMMA", "LDGSTS", "FHFMA"]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
kern, cur, i = OrderedDict(), None, 0
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = re.sub(r"\(.*", "", names[i].replace("(int)", "").replace("(bool)", "")); i += 1
        kern[cur] = Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and cur:
        op = m.group(1)
        kern[cur]["total"] += 1
        for c in COLS:
            if op == c or op.startswith(c + "."):
                kern[cur][c] += 1
out = open(sys.argv[1], "w") if len(sys.argv) > 1 else sys.stdout
out.write("# SASS mnemonic census of rgbmanip_b200/libadapose_b200.so (cuobjdump -sass, sm_100a), per kernel\n"
          "# UTCHMMA = tcgen05.mma kind::f16, UTCQMMA = kind::f8f6f4, UTMALDG / UTMASTG = TMA tensor load / store, LDTM = tcgen05.ld, "
          "UTCBAR = tcgen05.commit, LDGSTS = cp.async, FHFMA = mixed-precision fma.rn.f32.f16\n")
out.write("kernel,total_instructions," + ",".join(COLS) + "\n")
tot = Counter()
for k, c in kern.items():
    out.write(f"{k},{c['total']}," + ",".join(str(c[x]) for x in COLS) + "\n")
    tot.update(c)
out.write(f"TOTAL,{tot['total']}," + ",".join(str(tot[x]) for x in COLS) + "\n")
