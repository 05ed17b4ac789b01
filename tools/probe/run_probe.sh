#!/bin/bash
mkdir -p gpurun_out
{ nproc; free -g; nvidia-smi; nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv; lscpu | head -20; } > gpurun_out/box_info.txt 2>&1
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 200 > gpurun_out/probe_clocks.csv &
SMI=$!
./tools/probe/fp64_probe > gpurun_out/fp64_probe.txt 2>&1
python tools/probe/dgemm_peak.py gpurun_out/fp64_peak.json > gpurun_out/dgemm_peak.txt 2>&1
kill $SMI
cat gpurun_out/fp64_probe.txt gpurun_out/dgemm_peak.txt
