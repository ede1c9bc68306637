#!/bin/bash
# round-2 opener on the GPU: tensor peaks, then the variant-slots A/B
mkdir -p gpurun_out/r2_open
timeout 120 python tools/probes/tf32_peak.py > gpurun_out/r2_open/tf32_peak.json 2> gpurun_out/r2_open/tf32_peak.err; cat gpurun_out/r2_open/tf32_peak.json
bash tools/variant_slots_ab.sh
