"""Text summary of an .ncu-rep (one block per captured launch): the numbers DESIGN.md / VERDICT cite.
usage: python scripts/ncu_summary.py report.ncu-rep [max_launches]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
maxn = int(sys.argv[2]) if len(sys.argv) > 2 else 16
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
want = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid CTAs"), ("launch__block_size", "block threads"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe (HMMA) active %"),
    ("sm__inst_executed.sum", "warp instructions executed"),
    ("sm__inst_issued.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps / scheduler"),
    ("smsp__issue_active.avg.per_cycle_active", "issued warps / scheduler / cycle"),
]
print("# %s" % rep.split("/")[-1])
for n, r in enumerate(data[:maxn]):
    print("\n## launch %d: %s" % (n, r[col["Kernel Name"]][:150]))
    for key, label in want:
        if key in col and r[col[key]] != "":
            print("  %-34s %s %s" % (label, r[col[key]], units[col[key]]))
    pipes = []
    for h, i in col.items():
        if (h.startswith("sm__inst_executed_pipe_") or h.startswith("sm__pipe_")) and h.endswith("pct_of_peak_sustained_active"):
            try:
                pipes.append((float(r[i]), h.replace(".avg.pct_of_peak_sustained_active", "").replace(".sum.pct_of_peak_sustained_active", "")))
            except ValueError:
                pass
    pipes.sort(reverse=True)
    if pipes:
        print("  busiest pipes (% of peak, active):  " + ", ".join("%s %.1f" % (n2, v) for v, n2 in pipes[:8]))
    for key in ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
                "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg.per_second", "smsp__cycles_active.avg"):
        if key in col and r[col[key]] != "":
            print("  %-34s %s %s" % (key[:34], r[col[key]], units[col[key]]))
    stalls = []
    for h, i in col.items():
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
            try:
                stalls.append((float(r[i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
    stalls.sort(reverse=True)
    print("  stall reasons (warps per issue):   " + ", ".join("%s %.2f" % (n2, v) for v, n2 in stalls[:6]))
