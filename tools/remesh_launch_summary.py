"""Summarise an ncu launch list of one redistribution call (tools/remesh_prof.py under
`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv`):
per kernel of the LAST call, duration and DRAM bytes.  Usage: remesh_launch_summary.py file.csv calls"""
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 3
hdr = rows[0]
ki, mi, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
d = {}
for r in rows[1:]:
    d.setdefault((int(r[ii]), r[ki]), {})[r[mi]] = float(r[vi].replace(",", ""))
ids = sorted(d)
per_call = len(ids) // calls
total = 0.0
print(f"{'kernel':72s} {'us':>9s} {'DRAM rd MB':>11s} {'DRAM wr MB':>11s} {'GB/s':>8s}")
for k in ids[-per_call:]:
    v = d[k]
    us = v["gpu__time_duration.sum"] / 1e3
    rd, wr = v["dram__bytes_read.sum"] / 1e6, v["dram__bytes_write.sum"] / 1e6
    total += us
    name = k[1].replace("cvtx::remesh::<unnamed>::", "").replace("void ", "")
    print(f"{name[:72]:72s} {us:9.1f} {rd:11.1f} {wr:11.1f} {(rd + wr) / us * 1e3:8.0f}")
print(f"{'total kernel time of one call':72s} {total:9.1f}")
