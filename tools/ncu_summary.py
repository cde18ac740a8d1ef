"""Condense `ncu -i X.ncu-rep --page raw --csv` into a per-kernel text summary for profiles/.
usage: python tools/ncu_summary.py raw.csv > profiles/name.txt"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
stall = [h for h in hdr if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.per_cycle_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_atom.sum", "lts__t_sectors_srcunit_tex_op_red.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum"]
for r in rows[2:]:
    print("kernel:", r[idx["Kernel Name"]][:110])
    for w in want:
        if w in idx:
            print(f"  {w:70s} {r[idx[w]]:>16s} {units[idx[w]]}")
    vals = [(float(r[idx[n]].replace(",", "")) if r[idx[n]] not in ("", "n/a") else 0.0, n) for n in stall]
    tot = sum(v for v, _ in vals) or 1
    vals.sort(reverse=True)
    print("  stall reasons (pc sampling):", ", ".join(f"{n.split('stalled_')[1]} {100 * v / tot:.0f}%" for v, n in vals[:6]))
    print()
