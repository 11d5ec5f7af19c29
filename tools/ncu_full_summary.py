"""Key metrics of every launch in an `ncu --set full` report -> a small text table for profiles/ (run where ncu is
installed; no GPU needed).

    python tools/ncu_full_summary.py gpurun_out/x.ncu-rep profiles/r2_ncu_full_x.summary.txt "what was captured"
"""
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("launch__grid_size", "grid"),
    ("launch__registers_per_thread", "regs"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_%"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64_%"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu_%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_%"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "smem_ld_conflicts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum", "smem_st_conflicts"),
    ("sm__cycles_active.avg", "sm_cycles"),
    ("smsp__inst_executed.sum", "warp_insts"),
]


def main(rep, out, title=""):
    if rep.endswith(".csv"):     # the raw page already exported on the GPU box (ncu -i x.ncu-rep --page raw --csv)
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [(m, short) for m, short in METRICS if m in idx]
    lines = [f"# {title}", f"# source: ncu --set full --clock-control none ({rep}); one row per captured launch",
             "# " + " | ".join(f"{short} [{units[idx[m]]}]" for m, short in cols) + " | kernel"]
    for r in data:
        name = r[idx["Kernel Name"]] if "Kernel Name" in idx else ""
        lines.append(" | ".join(r[idx[m]] for m, _ in cols) + " | " + name[:70])
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:8]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
