"""Key figures of `ncu --set full` reports (one block per captured launch): python tools/summarize_ncu_full.py a.ncu-rep [b.ncu-rep ...]"""
import csv, subprocess, sys
KEYS = [("gpu__time_duration.sum", "duration"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "registers/thread"), ("launch__occupancy_limit_registers", "CTAs/SM limit: registers"),
        ("launch__occupancy_limit_shared_mem", "CTAs/SM limit: shared memory"), ("smsp__inst_executed.sum", "warp instructions executed"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (ex2) pipe %"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
        ("lts__t_sector_hit_rate.pct", "L2 hit rate %"), ("sm__cycles_active.avg", "SM active cycles (avg)"), ("sm__cycles_elapsed.max", "SM elapsed cycles")]
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    head, units = rows[0], rows[1]
    for row in rows[2:]:
        d, u = dict(zip(head, row)), dict(zip(head, units))
        print(f"== {d['Kernel Name'][:110]}   [{rep.split('/')[-1]}]")
        for k, label in KEYS:
            if d.get(k) not in (None, ""):
                print(f"   {label:34s} {d[k]} {u.get(k, '')}")
        dur = float(d["gpu__time_duration.sum"].replace(",", ""))
        dur_s = dur * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3}.get(u["gpu__time_duration.sum"], 1e-6)
        stalls = [(k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), float(v.replace(",", "")))
                  for k, v in d.items() if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio") and v not in ("", "n/a")]
        print("   top stalls (warps per issue)       " + ", ".join(f"{k} {v:.2f}" for k, v in sorted(stalls, key=lambda x: -x[1])[:6]))
