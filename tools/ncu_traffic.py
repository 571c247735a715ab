"""profiles/dominant_kernel_traffic.json from an `ncu --set full` raw csv: DRAM bytes (read + write) per launch of the
dominant kernel (the channel-MLP fc1 contraction = gemm_tc16_kernel<2, 1, 1, 0> launches of the capture), which bench.py
reports as roofline.traffic.

    python tools/ncu_traffic.py gpurun_out/r02y_fwd.raw.csv "gemm_tc16_kernel<2, 1, 1, 0>" profiles/r02y_ncu_fwd.raw.csv"""
import csv
import json
import os
import sys

path, pat = sys.argv[1], sys.argv[2]
src = sys.argv[3] if len(sys.argv) > 3 else path
rows = list(csv.reader(open(path)))
hdr, units, data = rows[0], rows[1], rows[2:]
ki = hdr.index("Kernel Name")


def col(name, r):
    i = hdr.index(name)
    v = float(r[i].replace(",", ""))
    return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units[i], 1.0)


max_us = float(sys.argv[4]) if len(sys.argv) > 4 else 1e30     # e.g. 60: leave out the ConvTranspose contraction (same instantiation, 80 us)
def _us(r):
    i = hdr.index("gpu__time_duration.sum")
    return float(r[i].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(units[i], 1e-3)
sel = [r for r in data if pat in r[ki] and _us(r) < max_us]
if not sel:
    raise SystemExit(f"no launch matching {pat!r} in {path}")
tot = [col("dram__bytes_read.sum", r) + col("dram__bytes_write.sum", r) for r in sel]
out = {"kernel": pat, "launches": len(sel), "dram_bytes_per_launch": sum(tot) / len(tot),
       "dram_read_bytes_per_launch": sum(col("dram__bytes_read.sum", r) for r in sel) / len(sel),
       "dram_write_bytes_per_launch": sum(col("dram__bytes_write.sum", r) for r in sel) / len(sel),
       "source": f"ncu --set full --clock-control none, {src} ({len(sel)} launches of one DPOT-S B=32 forward)"}
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
json.dump(out, open(os.path.join(root, "profiles", "dominant_kernel_traffic.json"), "w"), indent=1)
print(json.dumps(out))
