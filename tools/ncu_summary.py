"""Summarise an .ncu-rep (raw page) into the handful of numbers DESIGN.md / profiles/ quote.
usage: python tools/ncu_summary.py report.ncu-rep [more.ncu-rep ...]"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__cycles_active.avg", "sm__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active",
    "sm__inst_issued.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_lsu.sum",
    "sm__inst_executed_pipe_uniform.sum", "sm__inst_executed_pipe_cbu.sum", "sm__inst_executed_pipe_adu.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed_op_shared_ld.sum",
    "smsp__inst_executed_op_shared_st.sum", "smsp__inst_executed_op_global_st.sum",
]


def main():
    for path in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            print(f"== {path}: {r[hdr.index('Kernel Name')][:90]}")
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    print(f"  {k:70s} {r[i]:>18s} {units[i]}")
            stalls = [(float(r[i].replace(',', '')), h) for i, h in enumerate(hdr)
                      if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio") and r[i]]
            if not stalls:
                stalls = [(float(r[i].replace(',', '')), h) for i, h in enumerate(hdr)
                          if "issue_stalled" in h and h.endswith(".pct") and r[i]]
            for v, h in sorted(stalls, reverse=True)[:9]:
                print(f"  stall {h:75s} {v:10.3f}")


if __name__ == "__main__":
    main()
