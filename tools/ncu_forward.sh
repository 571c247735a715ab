#!/bin/bash
# Full ncu capture (--set full, source-level) of the library kernels of ONE DPOT-S forward (B=32), third forward of
# tools/one_forward.py (26 pack launches + 51 per forward match the filter): head = PatchEmbed, folded time aggregation,
# block 0 (7 kernels); tail = split, ConvTranspose GEMM, out tail.
# Output: gpurun_out/$1_{head,tail}.ncu-rep + raw CSVs (read here with ncu -i ... --page raw --csv).
name=${1:-fwd_full}
mkdir -p gpurun_out
RX='gemm_tc16|afno_fft|patch_embed|out_tail|split_f16|gn_finalize|spatial_mean|gemm_skinny'
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$RX" -s 128 -c 9 \
    -o gpurun_out/${name}_head -f python tools/one_forward.py S 32 > gpurun_out/${name}_head.log 2>&1
echo "ncu head exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$RX" -s 176 -c 3 \
    -o gpurun_out/${name}_tail -f python tools/one_forward.py S 32 > gpurun_out/${name}_tail.log 2>&1
echo "ncu tail exit $?"
for p in head tail; do ncu -i gpurun_out/${name}_$p.ncu-rep --page raw --csv > gpurun_out/${name}_$p.raw.csv 2>/dev/null; done
ls -la gpurun_out/; du -sh gpurun_out
