"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: the launches of the last
bench step (from the last absmax_kernel on) and per-kernel shares.  usage: launch_summary.py file.csv"""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ik, iv, ig = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
data = rows[1:]
start = [i for i, r in enumerate(data) if "absmax" in r[ik]][-1]
tot = 0.0
share = collections.OrderedDict()
def short(n):
    n = n.replace("fb200::<unnamed>::", "").replace("void ", "")
    return n.split("(")[0][:44]
for r in data[start:]:
    t = float(r[iv].replace(",", "")) / 1e3  # ns -> us
    tot += t
    share[short(r[ik])] = share.get(short(r[ik]), 0.0) + t
    print("%-46s %10.1f us  grid %s" % (short(r[ik]), t, r[ig]))
print("# total %.2f ms" % (tot / 1e3))
for k, v in sorted(share.items(), key=lambda kv: -kv[1]):
    print("# share %-44s %8.2f ms  %5.1f %%" % (k, v / 1e3, 100 * v / tot))
