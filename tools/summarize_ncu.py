"""Summarises an .ncu-rep (read here on the CPU box with `ncu -i`) into a small text table for profiles/.

    python tools/summarize_ncu.py gpurun_out/prof_1m_r01.ncu-rep > profiles/ncu_1m_r01.txt
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_allocated", "smem/block"),
    ("launch__occupancy_limit_registers", "occ limit (regs) blocks/SM"),
    ("launch__occupancy_limit_shared_mem", "occ limit (smem) blocks/SM"),
    ("launch__waves_per_multiprocessor", "waves/SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active % of peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__inst_executed.avg.per_cycle_elapsed", "IPC per SM"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe active %"),
    ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "FMA-heavy pipe active %"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe active %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("sm__inst_executed_pipe_tex.avg.pct_of_peak_sustained_active", "TEX pipe %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("dram__bytes_read.sum.per_second", "DRAM read rate"),
    ("dram__bytes_write.sum.per_second", "DRAM write rate"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("smsp__average_warp_latency_per_inst_issued.ratio", "warp cycles per issued inst"),
]
STALL_PREFIX = "smsp__average_warps_issue_stalled_"


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    print("# %s  (ncu --set full --clock-control none; one row per profiled launch)" % rep)
    for r in data:
        name = r[col["Kernel Name"]]
        print("\n== %s" % name[:110])
        for key, label in KEYS:
            if key in col and r[col[key]] != "":
                print("  %-34s %14s %s" % (label, r[col[key]], units[col[key]]))
        stalls = []
        for h, i in col.items():
            if h.startswith(STALL_PREFIX) and h.endswith("_per_issue_active.ratio") and r[i] not in ("", "0"):
                stalls.append((float(r[i]), h[len(STALL_PREFIX):-len("_per_issue_active.ratio")]))
        stalls.sort(reverse=True)
        print("  stall reasons (warps per issue-active cycle): " + ", ".join("%s %.2f" % (n, v) for v, n in stalls[:8]))


if __name__ == "__main__":
    main()
