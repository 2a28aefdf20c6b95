"""Aggregate an ncu --csv launch list (gpu__time_duration.sum) by kernel name: count, total, mean."""
import csv, sys, collections, re
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]; ik = hdr.index("Kernel Name"); iv = hdr.index("Metric Value"); iu = hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ik]).split("::")[-1]
    v = float(r[iv].replace(",", "")); v = v / 1e3 if r[iu] in ("ns", "nsecond") else v
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:40s} {n:6d} launches {t:12.1f} us total {t/n:10.2f} us mean {100*t/tot:5.1f} %")
print(f"{'total':40s} {sum(a[0] for a in agg.values()):6d} launches {tot:12.1f} us")
