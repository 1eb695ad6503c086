#!/bin/bash
# One GPU-box visit: parity tests, bench, ncu launch list (same command as the bench) and --set full / in-situ DRAM captures of one
# whole-chip ResNet-50 batch-32 encode.  Everything lands in gpurun_out/ (scratch, <= 64 MiB); tools/ncu_summary.py turns it into profiles/.
set -u
TAG=${1:-r2}
PER_ENCODE=${2:-53}   # kernel launches of one ResNet-50 batch-32 encode (boundary + 49 conv + pool + gap + fc)
CONV_PER_ENCODE=${3:-50}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$TAG.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
cut -c1-400 gpurun_out/bench_$TAG.json
cp gpurun_out/bench_profile_n1.json gpurun_out/bench_profile_$TAG.json 2>/dev/null
timeout 600 python tools/conv_layers.py > gpurun_out/conv_layers_$TAG.jsonl 2>&1
timeout 600 python tools/model_latency.py > gpurun_out/model_latency_$TAG.json 2> gpurun_out/model_latency_$TAG.err
# launch list of the bench command itself (cold-cache, serialised: compare shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/bench_under_ncu_$TAG.log 2>&1; echo "ncu list rc=$?"
# one whole-chip encode (graph replay off so that every kernel is a plain launch), --set full, raw page exported here
timeout 900 ncu --set full --clock-control none --launch-skip $((3 * PER_ENCODE)) -c $PER_ENCODE -f -o /tmp/encode_$TAG \
    python tools/one_encode.py 32 1 5 0 > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
ncu -i /tmp/encode_$TAG.ncu-rep --page raw --csv > gpurun_out/encode_${TAG}_raw.csv 2>/dev/null
# the same encode with warm caches (no flush between kernels): DRAM bytes each kernel really moves inside the step
timeout 900 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
    --launch-skip $((3 * PER_ENCODE)) -c $PER_ENCODE --csv --log-file gpurun_out/dram_insitu_$TAG.csv python tools/one_encode.py 32 1 5 0 > /dev/null 2>&1; echo "ncu insitu rc=$?"
# four conv launches with source correlation
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_ --launch-skip $((3 * CONV_PER_ENCODE + 20)) -c 4 -f -o gpurun_out/conv_src_$TAG \
    python tools/one_encode.py 32 1 5 0 > /dev/null 2>&1; echo "ncu src rc=$?"
ls -la gpurun_out | tail -12; du -sh gpurun_out
