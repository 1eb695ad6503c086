timeout 600 python tools/model_latency.py 2>&1 | tail -9
