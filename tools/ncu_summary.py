#!/usr/bin/env python
"""Condense an .ncu-rep (ncu --set full) into one CSV row per profiled launch with the metrics bench.py's roofline and
DESIGN.md quote: duration, DRAM bytes read/written (-> `traffic`), DRAM / L2 / tensor-pipe utilisation, registers.
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_<name>.csv"""
import csv
import subprocess
import sys

WANT = [
    ("Kernel Name", "kernel"), ("Grid Size", "grid"), ("Block Size", "block"),
    ("gpu__time_duration.sum", "duration_us"),
    ("dram__bytes_read.sum", "dram_read_MB"), ("dram__bytes_write.sum", "dram_write_MB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "l2_to_sm_read_MB"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_active_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_dynamic", "dyn_smem"),
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [(hdr.index(k), n) for k, n in WANT if k in hdr]
    w = csv.writer(sys.stdout)
    # byte counters come back in whatever unit ncu picked per column (byte / Kbyte / Mbyte / Gbyte): the *_MB columns are always MB
    scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}

    def cell(r, i, n):
        if n == "kernel":
            return r[i].split("(")[0]
        if n.endswith("_MB") and units[i] in scale:
            try:
                return "%.3f" % (float(r[i].replace(",", "")) * scale[units[i]])
            except ValueError:
                return r[i]
        return r[i]

    w.writerow([n + ("" if not units[i] or n in ("kernel", "grid", "block") or n.endswith("_MB") else "[" + units[i] + "]") for i, n in idx])
    for r in rows[2:]:
        w.writerow([cell(r, i, n) for i, n in idx])


if __name__ == "__main__":
    main(sys.argv[1])
