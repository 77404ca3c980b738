"""Summarise .ncu-rep captures (ncu --set full) into the table profiles/ keeps: duration, DRAM bytes, DRAM %,
tensor pipe %, occupancy, registers, top stall reasons."""
import csv, subprocess, sys, io
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
for path in sys.argv[1:]:
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = csv.reader(io.StringIO(raw)); hdr = next(rd); units = next(rd)
    print(f"== {path}")
    for r in rd:
        d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
        name = d["Kernel Name"].split("(")[0][-60:]
        print(f"  {name}  grid={d.get('launch__grid_size')} block={d.get('launch__block_size')} regs={d.get('launch__registers_per_thread')}")
        rdv, wrv = float(d["dram__bytes_read.sum"]), float(d["dram__bytes_write.sum"])
        print(f"    time {d['gpu__time_duration.sum']} {u['gpu__time_duration.sum']}; DRAM read {rdv:.2f} + write {wrv:.2f} {u['dram__bytes_read.sum']}"
              f" ({d['gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']} % of peak); L2 hit {d.get('lts__t_sector_hit_rate.pct','?')} %")
        print(f"    tensor pipe {d.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','?')} %; warps active "
              f"{d['sm__warps_active.avg.pct_of_peak_sustained_active']} %; issue active {d.get('smsp__issue_active.avg.pct_of_peak_sustained_active','?')} %")
        stalls = sorted(((float(v), k) for k, v in d.items() if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and v not in ("", "n/a")), reverse=True)[:3]
        print("    top stalls: " + ", ".join(f"{k.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio','')}={v:.2f}" for v, k in stalls))
