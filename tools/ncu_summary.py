"""Condense an .ncu-rep into the few per-kernel numbers DESIGN.md / bench.py cite (run where ncu is installed).
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_<name>.txt
"""
import csv
import io
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
           "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
           "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size", "sm__cycles_elapsed.max"]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print("# source:", path)
    seen = {}
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        name = r[col["Kernel Name"]].split("(")[0]
        seen.setdefault(name, []).append(r)
    for name, lst in seen.items():
        print("\n== %s   (%d launches captured)" % (name, len(lst)))
        for m in METRICS:
            if m not in col:
                continue
            vals = []
            for r in lst:
                try:
                    vals.append(float(r[col[m]].replace(",", "")))
                except ValueError:
                    pass
            if vals:
                print("  %-78s min %-14.6g max %-14.6g [%s]" % (m, min(vals), max(vals), units[col[m]]))


if __name__ == "__main__":
    main(sys.argv[1])
