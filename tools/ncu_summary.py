"""Markdown table of the headline counters of every kernel in an .ncu-rep (read here, no GPU needed):
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep"""
import csv, io, subprocess, sys

WANT = [("gpu__time_duration.sum", "duration"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "MUFU (XU) pipe %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % (ncu peak)"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput %"),
        ("launch__registers_per_thread", "registers / thread"),
        ("launch__occupancy_limit_registers", "CTAs / SM (register limit)"),
        ("launch__occupancy_limit_shared_mem", "CTAs / SM (shared-memory limit)"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("sass__inst_executed_local_loads", "local loads"), ("sass__inst_executed_local_stores", "local stores"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard / issue"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall: barrier / issue"),
        ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall: MIO throttle / issue"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall: math pipe throttle / issue")]

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print(f"\n**`{r[hdr.index('Kernel Name')][:110]}`**\n\n| counter | value |\n|---|---|")
    for key, label in WANT:
        if key in hdr:
            v = r[hdr.index(key)]
            try:
                v = f"{float(v):,.2f}".rstrip("0").rstrip(".")
            except ValueError:
                pass
            print(f"| {label} (`{key}`) | {v} {units[hdr.index(key)]} |")
