#!/usr/bin/env python
"""Turn an .ncu-rep (brought back in gpurun_out/) into the short text summary kept under profiles/."""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__cycles_elapsed.max"]


def main(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = [f"# {rep}"]
    for r in rows[2:]:
        out.append(f"kernel: {r[hdr.index('Kernel Name')]}")
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                out.append(f"  {w:78s} {r[i]:>18s} {units[i]}")
        out.append("  warp stall reasons (warps per issue-active cycle, > 0.5):")
        for i, h in enumerate(hdr):
            if "average_warps_issue_stalled" in h and float(r[i] or 0) > 0.5:
                out.append(f"    {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):30s} {float(r[i]):6.2f}")
    print("\n".join(out))


if __name__ == "__main__":
    main(sys.argv[1])
