#!/bin/bash
# Round-2 session u (not a test): ncu --set full with source of the current build (3256 units).
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_emit|k_match|k_split|k_link|k_gather" -c 5 -o gpurun_out/r2u_full -f \
    python tests/prof_run.py 3256 > gpurun_out/r2u_ncu_full.log 2>&1
tail -3 gpurun_out/r2u_ncu_full.log
