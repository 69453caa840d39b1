"""ncu -i <rep> --page raw --csv  ->  compact per-kernel summary CSV/markdown for profiles/."""
import csv, subprocess, sys, collections
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
cols = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_read"),
        ("dram__bytes_write.sum", "dram_write"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
        ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("lts__t_sector_hit_rate.pct", "l2_hit_pct")]
cols = [(c, n) for c, n in cols if c in idx]
seen = collections.Counter()
with open(out, "w") as f:
    w = csv.writer(f)
    w.writerow([n + (f" [{units[idx[c]]}]" if units[idx[c]] else "") for c, n in cols])
    for r in data:
        name = r[idx["Kernel Name"]]
        key = name.split("(")[0]
        seen[key] += 1
        if seen[key] > 2:
            continue
        w.writerow([r[idx[c]][:90] if n == "kernel" else r[idx[c]] for c, n in cols])
print(open(out).read())
