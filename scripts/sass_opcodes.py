"""Opcode evidence from the built objects (no GPU needed): per kernel, how often the Blackwell-native
instructions appear in SASS.  usage: python scripts/sass_opcodes.py > profiles/rNN_sass_opcodes.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "histogan_b200", "build")
KEYS = ["UTCHMMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "UTCBAR", "SYNCS", "FFMA", "MUFU", "RED", "ATOM", "LDGSTS", "HMMA"]

print("# SASS opcode counts per kernel (cuobjdump -sass of histogan_b200/build/*.o, sm_100a)\n")
print("`UTCHMMA` = tcgen05.mma, `UTMALDG` / `UTMASTG` = TMA tensor load / store, `UBLKCP` = 1-D bulk copy (cp.async.bulk), `LDTM` = tcgen05.ld, "
      "`UTCBAR` = tcgen05.commit, `SYNCS` = mbarrier ops, `HMMA` would be the legacy mma.sync path (absent).\n")
print("| object | kernel | " + " | ".join(KEYS) + " | instructions |")
print("|---|---|" + "---|" * (len(KEYS) + 1))
for f in sorted(os.listdir(OBJ)):
    if not f.endswith(".o"):
        continue
    out = subprocess.run(["cuobjdump", "-sass", os.path.join(OBJ, f)], capture_output=True, text=True).stdout
    name, counts = None, None
    rows = []
    for line in out.splitlines():
        m = re.match(r"\s+Function : (\S+)", line)
        if m:
            if name:
                rows.append((name, counts))
            name, counts = m.group(1), collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and name:
            op = m.group(1)
            counts["_n"] += 1
            for k in KEYS:
                if op.startswith(k):
                    counts[k] += 1
    if name:
        rows.append((name, counts))
    for name, c in rows:
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        dem = re.sub(r"\(.*", "", dem).replace("void ", "")
        if c["_n"] < 40 and not any(c[k] for k in KEYS[:5]):
            continue
        print(f"| {f} | `{dem[:70]}` | " + " | ".join(str(c[k]) if c[k] else "" for k in KEYS) + f" | {c['_n']} |")
