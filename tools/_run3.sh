TAG=r1
timeout 600 python tools/conv_layers.py > gpurun_out/conv_layers_$TAG.jsonl 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu_$TAG.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --launch-skip 171 -c 57 -f -o /tmp/encode_$TAG \
    python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
ncu -i /tmp/encode_$TAG.ncu-rep --page raw --csv > gpurun_out/encode_${TAG}_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_igemm --launch-skip 162 -c 4 -f -o gpurun_out/conv_src_$TAG \
    python bench.py --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1; echo "ncu src rc=$?"
cp gpurun_out/bench_profile_n1.json gpurun_out/bench_profile_$TAG.json
ls -la gpurun_out | tail -20; du -sh gpurun_out
