"""ncu --csv launch list (gpu__time_duration.sum) -> per-kernel and per-class table (markdown).
usage: python scripts/summarize_launches.py <csv> [title]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
hdr = rows[hi]
kn, mv, mu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
data = []
for r in rows[hi + 1:]:
    if len(r) <= mv or not r[mv]:
        continue
    v = float(r[mv].replace(',', ''))
    v = v / 1e3 if r[mu] in ('ns', 'nsecond') else v * (1e3 if r[mu] in ('ms', 'msecond') else 1)     # -> us
    data.append((r[kn], v))


def cls(k):
    if 'conv_tf32' in k or 'conv_wgrad' in k or 'conv_finish' in k:
        return 'tcgen05 conv'
    if 'hist' in k or 'hellinger' in k:
        return 'histogram'
    if k.startswith('hg::') or k.startswith('void hg::'):
        return 'hand-written element-wise / pack / optimiser'
    if 'cutlass' in k or 'gemm' in k or 'gemv' in k or 'globalKernel' in k or 'splitK' in k or 'scal_kernel' in k:
        return 'cuBLAS'
    if 'nccl' in k.lower():
        return 'NCCL'
    return 'torch (at::)'


agg, by_cls, n_cls = collections.OrderedDict(), collections.Counter(), collections.Counter()
for k, d in data:
    c = cls(k)
    by_cls[c] += d
    n_cls[c] += 1
    k = re.sub(r'\(.*', '', k)[:88]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += d
tot = sum(by_cls.values())
print(f"# {sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]}\n")
print(f"launches listed: {len(data)}, summed duration {tot / 1e3:.2f} ms "
      f"(per-launch times under ncu are serialised and cold-cache: shares matter, not absolutes)\n")
print("| class | launches | total us | share |\n|---|---|---|---|")
for c, v in by_cls.most_common():
    print(f"| {c} | {n_cls[c]} | {v:.0f} | {100 * v / tot:.1f} % |")
print("\n| kernel | launches | total us | share |\n|---|---|---|---|")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
    print(f"| `{k}` | {v[0]} | {v[1]:.0f} | {100 * v[1] / tot:.1f} % |")
