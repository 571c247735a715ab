#!/bin/bash
# Full ncu capture (--set full, source-level) of ONE DPOT-S forward (B=32), final round-2 launch sequence: PatchEmbed,
# folded time aggregation, 6 x (fused AFNO mixer incl. GroupNorm-2, fc1, fc2), cls head (3), ConvTranspose contraction,
# output tail = 25 matching launches, delimited by cudaProfilerStart/Stop around the third forward of tools/one_forward.py.
name=${1:-fwd_full}
mkdir -p gpurun_out
RX='afno_fused_kernel|gemm_tc16_kernel|split_f16_gn_rows|patch_embed_mma|out_tail_tc|spatial_mean'
timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"$RX" -c 30 \
    -o gpurun_out/${name} -f python tools/one_forward.py S 32 > gpurun_out/${name}.log 2>&1
echo "ncu exit $?"
ncu -i gpurun_out/${name}.ncu-rep --page raw --csv > gpurun_out/${name}.raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/${name}.raw.csv
