"""gpurun_out/parity_records.jsonl (written by the -m gpu tests through tests/parity.record)
-> a markdown table of the MEASURED parity errors.  usage: python scripts/parity_table.py > profiles/rNN_parity_table.md"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "parity_records.jsonl")
rows = {}
for line in open(path):
    r = json.loads(line)
    rows[r.pop("check")] = r            # last record of a check wins
print("# measured parity errors (B200, `pytest tests -m gpu`; tests/parity.record)\n")
print("| check | measured |\n|---|---|")
for k in sorted(rows):
    vals = ", ".join(f"{n} {v:.2e}" if isinstance(v, float) else f"{n} {v}" for n, v in rows[k].items())
    print(f"| `{k}` | {vals} |")
