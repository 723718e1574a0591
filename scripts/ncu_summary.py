#!/usr/bin/env python
"""Summarise an .ncu-rep: headline metrics + top stall locations by warp role (exec count)."""
import csv
import io
import subprocess
import sys


def raw_all(path):
    """One dict per profiled launch."""
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(out)))
    return [dict(zip(r[0], row)) for row in r[2:]]


def source_sections(path):
    """The source page lists the profiled launches one after the other, each with its own header row."""
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    heads = [i for i, r in enumerate(rows) if "# Samples" in r]
    return [(rows[h], rows[h + 1: (heads[j + 1] - 1 if j + 1 < len(heads) else len(rows))]) for j, h in enumerate(heads)]


KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum", "sm__cycles_elapsed.avg",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]


def main(path, top=25):
    launches = raw_all(path)
    sections = source_sections(path)
    for li, m in enumerate(launches):
        print(f"==== launch {li}: {m.get('Kernel Name', '?')[:100]}  grid {m.get('Grid Size', '?')} block {m.get('Block Size', '?')}")
        for k in KEYS:
            if k in m:
                print(f"{k} = {m[k]}")
        if len(sections) != len(launches) or li >= len(sections):
            continue
        hdr, rows = sections[li]
        si = hdr.index("# Samples")
        ei = hdr.index("Instructions Executed") if "Instructions Executed" in hdr else None
        stall = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith("stall") and "Not Issued" not in h]
        data = []
        for idx, r in enumerate(rows):
            if len(r) <= si or not (r[si] or "0").replace(",", "").isdigit():
                continue
            s_, e = int(r[si] or 0), int(r[ei] or 0) if ei is not None and len(r) > ei and (r[ei] or "0").isdigit() else 0
            reasons = {h: int(r[i]) for i, h in stall if i < len(r) and (r[i] or "0").isdigit() and int(r[i] or 0) > 0}
            data.append((s_, e, idx, r[1].strip()[:64], reasons))
        tot = sum(d[0] for d in data) or 1
        print("total samples", tot)
        for d in sorted(data, reverse=True)[:top]:
            print(f"{d[0]:6d} {100 * d[0] / tot:5.1f}% exec={d[1]:8d} #{d[2]:5d} {d[3]:64s} {d[4]}")


def stalls_only(path, top=25):
    """When the source page does not split per launch: one table per source section."""
    for si_, (hdr, rows) in enumerate(source_sections(path)):
        si = hdr.index("# Samples")
        stall = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith("stall") and "Not Issued" not in h]
        data = []
        for idx, r in enumerate(rows):
            if len(r) <= si or not (r[si] or "0").isdigit():
                continue
            reasons = {h: int(r[i]) for i, h in stall if i < len(r) and (r[i] or "0").isdigit() and int(r[i] or 0) > 0}
            data.append((int(r[si] or 0), idx, r[1].strip()[:64], reasons))
        tot = sum(d[0] for d in data) or 1
        print(f"==== source section {si_}: total samples {tot}")
        for d in sorted(data, reverse=True)[:top]:
            print(f"{d[0]:6d} {100 * d[0] / tot:5.1f}% #{d[1]:5d} {d[2]:64s} {d[3]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
    if len(source_sections(sys.argv[1])) != len(raw_all(sys.argv[1])):
        stalls_only(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
