#!/bin/bash
# One GPU-box visit: parity tests, bench, ncu launch list (same command as the bench) and one --set full capture of a whole
# ResNet-50 batch-32 encode.  Everything lands in gpurun_out/ (scratch, <= 64 MiB); tools/ncu_summary.py turns it into profiles/.
set -u
TAG=${1:-r1}
PER_ENCODE=${2:-53}   # kernel launches of one ResNet-50 batch-32 encode (boundary + 49 conv + pool + gap + fc)
CONV_PER_ENCODE=${3:-50}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$TAG.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
cat gpurun_out/bench_$TAG.json
cp gpurun_out/bench_profile_n1.json gpurun_out/bench_profile_$TAG.json 2>/dev/null
timeout 600 python tools/conv_layers.py > gpurun_out/conv_layers_$TAG.jsonl 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu_$TAG.log 2>&1; echo "ncu list rc=$?"
# whole encode, --set full, raw page exported here (the .ncu-rep of ~55 launches is > 64 MiB and would not travel back)
timeout 900 ncu --set full --clock-control none --launch-skip $((3 * PER_ENCODE)) -c $PER_ENCODE -f -o /tmp/encode_$TAG \
    python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
ncu -i /tmp/encode_$TAG.ncu-rep --page raw --csv > gpurun_out/encode_${TAG}_raw.csv 2>/dev/null
# three representative conv launches with source correlation (3x3 im2col, 1x1 tiled + residual, stem)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_ --launch-skip $((3 * CONV_PER_ENCODE)) -c 4 -f -o gpurun_out/conv_src_$TAG \
    python bench.py --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1; echo "ncu src rc=$?"
ls -la gpurun_out | tail -20; du -sh gpurun_out
