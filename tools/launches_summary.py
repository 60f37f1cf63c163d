"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.  usage: launches_summary.py launches.csv "header comment" """
import collections, csv, re, sys

path, title = sys.argv[1:3]
rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("==")) if len(r) > 5]
h = rows[0]
ik, iv, iu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    v = float(r[iv].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu], 1.0)
    name = re.sub(r"^void |mopa::|\(.*$", "", r[ik])
    tot[name] += v; cnt[name] += 1
s = sum(tot.values())
print("# " + title)
print("kernel,launches,total_us,share")
for k, v in tot.most_common():
    print("%s,%d,%.1f,%.4f" % (k, cnt[k], v, v / s))
