#!/bin/bash
# Full ncu capture (--set full, source-level) of ONE DPOT-S forward (B=32) with the fused AFNO mixer: the third forward of
# tools/one_forward.py.  28 matching launches per forward: PatchEmbed, folded time aggregation, 6 x (fused mixer,
# GroupNorm-2 split, fc1, fc2), ConvTranspose contraction, output tail.  head = the first 10, tail = the last 3.
name=${1:-fwd_full}
mkdir -p gpurun_out
RX='afno_fused_kernel|gemm_tc16_kernel|split_f16_gn_rows|patch_embed_mma|out_tail_tc'
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$RX" -s 56 -c 10 \
    -o gpurun_out/${name}_head -f python tools/one_forward.py S 32 > gpurun_out/${name}_head.log 2>&1
echo "ncu head exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$RX" -s 81 -c 3 \
    -o gpurun_out/${name}_tail -f python tools/one_forward.py S 32 > gpurun_out/${name}_tail.log 2>&1
echo "ncu tail exit $?"
for p in head tail; do ncu -i gpurun_out/${name}_$p.ncu-rep --page raw --csv > gpurun_out/${name}_$p.raw.csv 2>/dev/null; python tools/ncu_summary.py gpurun_out/${name}_$p.raw.csv; done
