"""DRAM bytes per launch (mean over every launch in the report) from an `ncu --set full` capture -> a small JSON
that bench.py reads for `roofline.traffic`.
usage: python scripts/ncu_traffic2.py report.ncu-rep out.json "<what was captured>" """
import csv, io, json, os, subprocess, sys
rep, outp, desc = sys.argv[1], sys.argv[2], sys.argv[3]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics",
                      "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
ir, iw, it = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
launches = []
for r in data:
    launches.append({"kernel": r[hdr.index("Kernel Name")][:80], "grid": r[hdr.index("Grid Size")],
                     "dram_read": float(r[ir]) * mul[units[ir]], "dram_write": float(r[iw]) * mul[units[iw]], "time_" + units[it]: float(r[it])})
tot = sum(l["dram_read"] + l["dram_write"] for l in launches)
j = {"dram_bytes_per_launch": tot / max(1, len(launches)), "dram_bytes_total": tot, "launches_in_report": len(launches), "launches": launches,
     "source": "ncu --set full --clock-control none, %s (%s)" % (os.path.basename(rep), desc)}
json.dump(j, open(outp, "w"), indent=1)
print(json.dumps(j)[:300])
