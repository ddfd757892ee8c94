"""Summarise `ncu --page raw --csv` exports (one row per captured launch) into a markdown table.
  python tools/ncu_raw_summary.py gpurun_out/r02_k2_C5_raw.csv [...]"""
import csv
import sys

KEYS = [("gpu__time_duration.sum", "duration"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs"), ("launch__shared_mem_per_block_dynamic", "dyn smem"),
        ("launch__waves_per_multiprocessor", "waves/SM"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe %"),
        ("sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "DMMA %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
        ("sm__pipe_tensor_subpipe_tcgen05_cycles_active.avg.pct_of_peak_sustained_active", "tcgen05 %"),
        ("smsp__inst_executed.sum", "warp instructions"),
        ("sass__inst_executed_local_loads", "local loads"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long sb"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short sb"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math"),
        ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no-inst")]


def main(paths):
    for path in paths:
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        idx = {h: i for i, h in enumerate(hdr)}
        print("\n### %s\n" % path.split("/")[-1])
        print("| metric | " + " | ".join("launch %d" % i for i in range(len(rows) - 2)) + " |")
        print("|---|" + "---:|" * (len(rows) - 2))
        kn = idx.get("Kernel Name")
        if kn is not None:
            print("| kernel | " + " | ".join("`%s`" % r[kn].split("(")[0][-48:] for r in rows[2:]) + " |")
        for key, label in KEYS:
            if key in idx:
                i = idx[key]
                print("| %s (%s) | " % (label, units[i]) + " | ".join(r[i] for r in rows[2:]) + " |")


if __name__ == "__main__":
    main(sys.argv[1:])
