"""Reads an `ncu --set full` report of the flat_tc launches of ONE bench step and writes
profiles/flat_tc_traffic.json (DRAM bytes per launch = mean over the step's launches).
usage: python scripts/ncu_traffic.py gpurun_out/prof_flat_tc.ncu-rep <launches_per_step>"""
import csv, io, json, os, subprocess, sys
rep, per_step = sys.argv[1], int(sys.argv[2])
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics",
                      "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
ir, iw, it = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
launches = []
for r in data[:per_step]:
    rd = float(r[ir]) * mul[units[ir]]
    wr = float(r[iw]) * mul[units[iw]]
    launches.append({"grid": r[hdr.index("Grid Size")], "dram_read": rd, "dram_write": wr, "time_" + units[it]: float(r[it])})
assert len(launches) == per_step, "report holds %d launches, expected %d" % (len(launches), per_step)
tot = sum(l["dram_read"] + l["dram_write"] for l in launches)
j = {"dram_bytes_per_launch": tot / per_step, "dram_bytes_per_step": tot, "launches_per_step": per_step, "launches": launches,
     "source": "ncu --set full --clock-control none, %s (flat_tc_kernel launches of one bench.py step, N=10M d=128 nq=10k k=100)" % os.path.basename(rep)}
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
json.dump(j, open(os.path.join(root, "profiles", "flat_tc_traffic.json"), "w"), indent=1)
print(json.dumps(j)[:400])
