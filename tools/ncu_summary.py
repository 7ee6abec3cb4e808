"""Summarise one `ncu --set full` report (.ncu-rep) into the few numbers DESIGN.md / bench.py quote:
    python tools/ncu_summary.py gpurun_out/prof_X.ncu-rep > profiles/r02_ncu_X.txt
Reads the report with `ncu -i ... --page raw --csv` (no GPU needed)."""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers/thread"), ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("launch__occupancy_limit_shared_mem", "CTAs/SM (smem limit)"), ("sm__warps_active.avg.per_cycle_active", "warps active / SM"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/shared throughput %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots active %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "pipe ALU %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "pipe FMA %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "pipe LSU %"),
    ("sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "pipe TMEM %"),
    ("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor-memory cycles active %"),
    ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe active % (realtime)"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared wavefronts (LSU)"),
    ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "shared wavefronts (tensor core operand reads)"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "shared pipe % (LSU)"),
    ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "shared pipe % (tensor core)"),
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        get = lambda k: (r[next(i for i, h in enumerate(hdr) if h.endswith(k))], units[next(i for i, h in enumerate(hdr) if h.endswith(k))])
        print(f"kernel: {r[hdr.index('Kernel Name')]}")
        print(f"report: {path} (ncu --set full --clock-control none, one launch)")
        for k, name in KEYS:
            try:
                v, u = get(k)
                print(f"  {name:48s} {v} {u}")
            except StopIteration:
                pass
        stalls = []
        for i, h in enumerate(hdr):
            if "average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                try:
                    stalls.append((float(r[i]), h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")))
                except ValueError:
                    pass
        print("  warp stall reasons (warps stalled per issue-active cycle):")
        for v, n in sorted(stalls, reverse=True)[:8]:
            print(f"    {n:28s} {v:.2f}")


if __name__ == "__main__":
    main(sys.argv[1])
