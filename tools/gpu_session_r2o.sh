#!/bin/bash
# Round 2, session o: wide-warp-tile DGEMM configurations on the C2 assembly shapes.
mkdir -p gpurun_out
timeout 500 python tools/gemm_tune2.py > gpurun_out/r02_o_gemm_wide_warp_tiles.jsonl 2> gpurun_out/o_tmp.err
cut -c1-200 gpurun_out/r02_o_gemm_wide_warp_tiles.jsonl; tail -3 gpurun_out/o_tmp.err
