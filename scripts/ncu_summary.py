"""Key numbers of every kernel in an ncu report: duration, regs, occupancy, IPC, pipe utilisation, stall breakdown."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[0]
def col(name):
    return h.index(name) if name in h else None
want = [
    ("Kernel Name", "kernel"), ("gpu__time_duration.sum", "dur_us"), ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active%"),
    ("smsp__inst_executed.sum", "warp_instr"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active%"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu(MUFU)%"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma%"),
    ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "fmaheavy%"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "alu%"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
    ("l1tex__data_pipe_lsu_wavefronts.sum", "lsu_wavefronts"),
    ("l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active", "lsu_wb%"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1_datapipe%"),
    ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("sm__cycles_elapsed.max", "sm_cycles"),
]
for r in rows[2:]:
    print("-" * 100)
    for name, label in want:
        c = col(name)
        if c is not None:
            print(f"  {label:18s} {r[c]}")
    # stall reasons (per-issue cycles)
    st = []
    for i, n in enumerate(h):
        if n.startswith("smsp__average_warps_issue_stalled_") and n.endswith("_per_issue_active.ratio") or \
           (n.startswith("smsp__average_warp_latency_issue_stalled_") and n.endswith(".ratio")):
            try:
                st.append((float(r[i]), n.replace("smsp__average_warps_issue_stalled_", "").replace("smsp__average_warp_latency_issue_stalled_", "").replace("_per_issue_active.ratio", "").replace(".ratio", "")))
            except ValueError:
                pass
    st.sort(reverse=True)
    print("  stalls/issue:", ", ".join(f"{n} {v:.2f}" for v, n in st[:10]))
