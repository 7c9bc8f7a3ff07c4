#!/usr/bin/env python
"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name: launches, total time, share."""
import csv, sys, collections, re
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr, rows = rows[0], rows[1:]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
if len(sys.argv) >= 5 and sys.argv[2] == "--last-step":
    # summarize_launches.py list.csv --last-step KERNEL PER_STEP: keep the launches of the last step only, a step ending with the
    # PER_STEP-th launch of KERNEL (e.g. adam_kernel 2: the generator update that ends a GAN step)
    kern, per = sys.argv[3], int(sys.argv[4])
    ends = [i for i, r in enumerate(rows) if kern in r[ki]]
    assert len(ends) >= 2 * per and len(ends) % per == 0, (len(ends), per)
    rows = rows[ends[-per - 1] + 1:ends[-1] + 1]
agg = collections.OrderedDict()
for r in rows:
    v = float(r[vi].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1e-3)
    name = re.sub(r"\(.*", "", r[ki])
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print("kernel,launches,total_us,share")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%s,%d,%.1f,%.4f" % (k, a[0], a[1], a[1] / tot))
print("TOTAL,%d,%.1f,1.0" % (sum(a[0] for a in agg.values()), tot))
