timeout 90 python __graft_entry__.py --smoke 2>&1 | tail -6
echo "== golden/resnet4"
timeout 200 python -m pytest tests/test_gpu_models.py -x -q -k "golden or batch4" 2>&1 | tail -15
