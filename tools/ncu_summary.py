"""Summarise an ncu --set full report: python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep"""
import csv
import io
import re
import subprocess
import sys

WANT = [r"^gpu__time_duration\.sum$", r"^launch__(registers_per_thread|grid_size|block_size|occupancy_limit.*|waves_per_multiprocessor)$",
        r"^sm__throughput\.avg\.pct", r"^smsp__issue_active\.avg\.pct", r"^sm__inst_executed_pipe_(fma|fmaheavy|fmalite|alu|xu|lsu|fp64|uniform)\.avg\.pct_of_peak_sustained_active$",
        r"^sm__pipe_(fma|fmaheavy|fmalite|alu|xu|fp64)_cycles_active\.avg\.pct_of_peak_sustained_active$",
        r"^smsp__warps_(active|eligible)\.avg\.per_cycle_active$", r"^sm__warps_active\.avg\.pct_of_peak_sustained_active$",
        r"^smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio$", r"^dram__bytes_(read|write)\.sum$", r"^lts__t_bytes\.sum$",
        r"^smsp__inst_executed\.sum$", r"^sm__cycles_elapsed\.avg$", r"^smsp__sass_thread_inst_executed_op_(ffma|fmul|fadd|fp32)_pred_on\.sum$",
        r"^smsp__inst_executed_pipe_(fma|xu|lsu|alu|fp64|fmaheavy|fmalite)\.sum$", r"^l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum$"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        print("kernel:", d.get("Kernel Name"), " grid", d.get("Grid Size"), " block", d.get("Block Size"))
        for k in hdr:
            if any(re.search(w, k) for w in WANT):
                v = d[k]
                if k.startswith("smsp__average_warps_issue_stalled") and float(v.replace(",", "") or 0) < 0.02:
                    continue
                print(f"  {k:95s} {v:>18s} {units[hdr.index(k)]}")


if __name__ == "__main__":
    for p in sys.argv[1:]:
        main(p)
