python -m pytest tests/test_gpu_models.py -q -k batch32 2>&1 | tail -40
for l in s1_1x1_64_256 s1_3x3_64 s1_1x1_64_256_res s3_1x1_256_1024_res s3_3x3_256 s4_1x1_2048_512 stem; do SMELTER_CONV_TIMELINE=1 python tools/conv_layers.py $l 2>&1 | tail -3; done
