"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): launches, total ms, per launch, share."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if r and r[0] == "ID":
        hdr, start = r, i + 1
        break
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[start:]:
    if len(r) <= iv:
        continue
    name = r[ik].split("(")[0].replace("void ", "")
    v = float(r[iv].replace(",", ""))
    ms = v / 1e6 if r[iu].startswith("n") else v / 1e3 if r[iu].startswith("u") else v
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ms
tot = sum(a[1] for a in agg.values())
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 16]:
    print("%-50s %5d %10.3f ms %8.3f per launch %5.1f%%" % (k[:50], n, ms, ms / n, 100 * ms / tot))
print("total %.3f ms" % tot)
