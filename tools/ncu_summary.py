"""Summarise .ncu-rep captures (ncu --set full) into the table profiles/ keeps: duration, DRAM bytes, DRAM %,
tensor pipe %, occupancy, registers, top stall reasons.  Every metric is converted with ITS OWN unit column (ncu
picks byte / Kbyte / Mbyte / Gbyte per metric).
  python tools/ncu_summary.py a.ncu-rep [b.ncu-rep ...]            -> text
  python tools/ncu_summary.py --json out.json a.ncu-rep ...        -> also per-kernel records as JSON"""
import csv, io, json, subprocess, sys
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3,
         "msecond": 1e3, "nsecond": 1e-3, "second": 1e6, "s": 1e6}


def num(d, u, key):
    v = d.get(key, "")
    if v in ("", "n/a"):
        return None
    return float(v.replace(",", "")) * SCALE.get(u.get(key, ""), 1.0)


def main():
    args = sys.argv[1:]
    jpath = None
    if args and args[0] == "--json":
        jpath, args = args[1], args[2:]
    recs = []
    for path in args:
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rd = csv.reader(io.StringIO(raw)); hdr = next(rd); units = next(rd)
        print(f"== {path}")
        for r in rd:
            d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
            name = d["Kernel Name"].split("(")[0][-70:]
            t_us = num(d, u, "gpu__time_duration.sum")
            rd_b, wr_b = num(d, u, "dram__bytes_read.sum") or 0.0, num(d, u, "dram__bytes_write.sum") or 0.0
            pct = lambda k: d.get(k, "?")
            print(f"  {name}  grid={d.get('launch__grid_size')} block={d.get('launch__block_size')} regs={d.get('launch__registers_per_thread')}")
            print(f"    time {t_us:.2f} us; DRAM read {rd_b / 1e6:.2f} MB + write {wr_b / 1e6:.2f} MB = {(rd_b + wr_b) / 1e6:.2f} MB "
                  f"-> {(rd_b + wr_b) / (t_us * 1e-6) / 1e9:.0f} GB/s ({pct('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')} % of ncu's DRAM peak); "
                  f"L2 hit {pct('lts__t_sector_hit_rate.pct')} %")
            print(f"    tensor pipe {pct('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active')} %; warps active "
                  f"{pct('sm__warps_active.avg.pct_of_peak_sustained_active')} %; issue active {pct('smsp__issue_active.avg.pct_of_peak_sustained_active')} %")
            stalls = sorted(((float(v), k) for k, v in d.items() if k.startswith("smsp__average_warps_issue_stalled_") and
                             k.endswith("_per_issue_active.ratio") and v not in ("", "n/a")), reverse=True)[:3]
            print("    top stalls: " + ", ".join(f"{k.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio','')}={v:.2f}" for v, k in stalls))
            recs.append({"file": path, "kernel": name, "grid": d.get("launch__grid_size"), "time_us": t_us, "dram_read_bytes": rd_b,
                         "dram_write_bytes": wr_b, "dram_pct": pct("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                         "tensor_pipe_pct": pct("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                         "warps_active_pct": pct("sm__warps_active.avg.pct_of_peak_sustained_active")})
    if jpath:
        with open(jpath, "w") as f:
            json.dump(recs, f, indent=1)


if __name__ == "__main__":
    main()
