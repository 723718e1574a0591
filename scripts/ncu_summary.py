#!/usr/bin/env python
"""Summarise an .ncu-rep: headline metrics + top stall locations by warp role (exec count)."""
import csv
import io
import subprocess
import sys


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(out)))
    return dict(zip(r[0], r[2]))


def source(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[1], rows[2:]


def main(path, top=25):
    m = raw(path)
    keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
            "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum", "sm__cycles_elapsed.avg",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
    for k in keys:
        if k in m:
            print(f"{k} = {m[k]}")
    hdr, rows = source(path)
    si, ei = hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith("stall") and "Not Issued" not in h]
    data = []
    for idx, r in enumerate(rows):
        s, e = int(r[si] or 0), int(r[ei] or 0)
        reasons = {h: int(r[i] or 0) for i, h in stall if int(r[i] or 0) > 0}
        data.append((s, e, idx, r[1].strip()[:64], reasons))
    tot = sum(d[0] for d in data)
    print("total samples", tot)
    for d in sorted(data, reverse=True)[:top]:
        print(f"{d[0]:6d} {100 * d[0] / tot:5.1f}% exec={d[1]:8d} #{d[2]:5d} {d[3]:64s} {d[4]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
