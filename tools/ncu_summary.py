"""Condenses an .ncu-rep (ncu --set full) into the handful of numbers DESIGN.md and bench.py
quote: duration, DRAM bytes, tensor/FP64 pipe activity, registers, occupancy, stall reasons.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.txt"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__cycles_elapsed.max", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "lts__t_bytes.sum", "sm__sass_inst_executed_op_shared_ld.sum"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        rec = dict(zip(hdr, zip(units, vals)))
        print("kernel:", rec.get("Kernel Name", ("", "?"))[1][:160])
        print("grid/block:", rec.get("Grid Size", ("", "?"))[1], rec.get("Block Size", ("", "?"))[1])
        for k in KEYS:
            if k in rec:
                print(f"  {k:86s} {rec[k][1]} {rec[k][0]}")
        for k in hdr:
            if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and float(rec[k][1] or 0) >= 0.05:
                print(f"  {k:86s} {rec[k][1]}")
        print()


if __name__ == "__main__":
    main(sys.argv[1])
