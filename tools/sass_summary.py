"""cuobjdump -sass opcode summary of libsealb200.so per kernel: which kernels issue tcgen05 MMAs (UTCHMMA*), TMA loads
(UTMALDG*), TMEM loads (LDTM*), griddepcontrol (ACQBULK / PDL), etc.   python tools/sass_summary.py [lib] > profiles/rNN_sass_summary.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "seal_b200/libsealb200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["c++filt", "-p", n], capture_output=True, text=True).stdout.strip()
kern = None
ops = collections.defaultdict(collections.Counter)
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1); continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        ops[kern][m.group(1)] += 1
KEY = ("UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "SYNCS", "HMMA", "IMMA", "FFMA", "POPC", "LDG", "STG", "LDS", "STS", "ATOM", "RED",
       "LDL", "STL", "ACQBULK", "BAR", "MUFU")
print(f"# {lib}: SASS opcode counts per kernel (sm_100a); columns = mnemonic prefixes")
for k in sorted(ops):
    c = ops[k]
    tot = sum(c.values())
    name = demangle(k)[:150]
    if "cub::" in name:
        continue
    cols = []
    for key in KEY:
        n = sum(v for op, v in c.items() if op.startswith(key))
        if n:
            cols.append(f"{key}={n}")
    extra = sorted({op for op in c if op.startswith(("UTC", "UTM", "LDTM", "UGETNEXT", "ELECT", "UCGABAR", "PREEXIT", "ACQ"))})
    print(f"{name}\n    instrs={tot} " + " ".join(cols) + ("\n    blackwell ops: " + ", ".join(extra) if extra else ""))
