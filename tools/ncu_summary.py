import csv, subprocess, sys
rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum", "sm__cycles_active.avg"]
with open(out, "w") as f:
    f.write(title + "\n\n| metric | value | unit |\n|---|---:|---|\n")
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            f.write(f"| {w} | {vals[i]} | {units[i]} |\n")
    for i, h in enumerate(hdr):
        if "issue_stalled" in h and "per_issue_active" in h and "not_issued" not in h:
            try:
                if float(vals[i]) >= 0.05:
                    f.write(f"| {h} | {float(vals[i]):.3f} | {units[i]} |\n")
            except ValueError:
                pass
print(open(out).read())
