export SMELTER_CONV_INSTRUMENT=1
for l in s1_1x1_64_64 s1_3x3_64 s3_1x1_1024_256 s3_3x3_256 s4_1x1_2048_512; do SMELTER_CONV_TIMELINE=1 python tools/conv_layers.py $l 2>&1 | tail -2; SMELTER_NO_PDL=1 SMELTER_CONV_TIMELINE=1 python tools/conv_layers.py $l 2>&1 | tail -2; done
