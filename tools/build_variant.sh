#!/bin/bash
# Build the library with compile-time experiment switches into smelter_b200/_variants/NAME.so (travels to the GPU box; load it with
# SMELTER_LIB_PATH=smelter_b200/_variants/NAME.so), then restore the default build.
# usage: tools/build_variant.sh NAME VAR=VAL ...
set -e
name=$1; shift
mkdir -p smelter_b200/_variants
env "$@" python -m smelter_b200.build --force > /dev/null
cp smelter_b200/libsmelter_b200.so smelter_b200/_variants/$name.so
echo "built smelter_b200/_variants/$name.so ($*)"
