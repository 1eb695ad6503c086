#!/bin/bash
# elementwise kernels: parity, HBM fractions (CUDA events), and the ncu evidence (DRAM bytes + duration per kernel)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_elementwise.py -q 2>&1 | tail -4
timeout 400 python tools/ew_bench.py > gpurun_out/ew_bench_r2.txt 2>&1; cat gpurun_out/ew_bench_r2.txt
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed --clock-control none \
   --csv --log-file gpurun_out/ew_ncu_r2.csv python tools/ew_bench.py > gpurun_out/ew_under_ncu.txt 2>&1; echo "ncu rc=$?"; wc -l gpurun_out/ew_ncu_r2.csv
