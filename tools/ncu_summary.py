"""Summarise `ncu --page raw --csv` output: one line per captured launch with the metrics that locate the bound.
    python tools/ncu_summary.py gpurun_out/x.raw.csv [more.csv ...]"""
import csv, re, sys
COLS = [
    ("gpu__time_duration.sum", "us", 1e-3),
    ("dram__bytes_read.sum", "rdMB", None),
    ("dram__bytes_write.sum", "wrMB", None),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%", 1),
    ("lts__t_bytes.sum", "L2MB", None),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%", 1),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tens%", 1),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%", 1),
    ("sm__issue_active.avg.pct_of_peak_sustained_active", "issue%", 1),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smemwf%", 1),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%", 1),
    ("launch__registers_per_thread", "regs", 1),
]
def tomb(v, unit):
    f = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1e-6)
    return v * f
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    print(f"# {path}")
    print(f"{'kernel':34s}" + "".join(f"{c[1]:>9s}" for c in COLS))
    for r in data:
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("dpot::<unnamed>::", "")[:33]
        out = f"{name:34s}"
        for m, lab, sc in COLS:
            if m not in hdr:
                out += f"{'-':>9s}"; continue
            i = hdr.index(m)
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:
                out += f"{'-':>9s}"; continue
            if sc is None:
                v = tomb(v, units[i])
            elif lab == "us":
                v = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(units[i], 1e-3)
            out += f"{v:9.1f}"
        print(out)
