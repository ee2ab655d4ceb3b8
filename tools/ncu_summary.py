"""Key metrics of every launch in an .ncu-rep (read with `ncu -i ... --page raw --csv`), as a markdown table.
  python tools/ncu_summary.py gpurun_out/r1b_hot.ncu-rep > profiles/r1_hot_kernels.md
"""
import csv, io, subprocess, sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe active %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem LSU wavefronts %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("sm__cycles_elapsed.avg", "SM cycles"),
]

def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"source: `{rep}` (ncu --set full --clock-control none; per-launch values are cold-cache and serialised)\n")
    names = [r[col["Kernel Name"]].split("(")[0].split("::")[-1] for r in rows[2:]]
    print("| metric | " + " | ".join(f"{i}: {n}" for i, n in enumerate(names)) + " |")
    print("|---|" + "---:|" * len(names))
    for key, label in WANT:
        if key not in col:
            continue
        i = col[key]
        print(f"| {label} [{units[i]}] | " + " | ".join(r[i] for r in rows[2:]) + " |")

if __name__ == "__main__":
    main()
