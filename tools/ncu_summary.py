"""Summarise an `ncu --set full` report (via `ncu -i REP --page raw --csv`) into the text kept under profiles/."""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"]


def main(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[h.index("Kernel Name")]
        print(f"## {name}   grid {r[h.index('Grid Size')]} block {r[h.index('Block Size')]}")
        for k in KEYS:
            if k in h:
                print(f"{k:75s} {r[h.index(k)]} {units[h.index(k)]}")
        st = {k: float(r[i]) for i, k in enumerate(h) if k.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in k and r[i]}
        tot = sum(st.values())
        if tot > 0:
            for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:7]:
                print(f"pc-sampling stall {k.replace('smsp__pcsamp_warps_issue_stalled_', ''):56s} {100 * v / tot:5.1f} %")
        print()


if __name__ == "__main__":
    main(sys.argv[1])
