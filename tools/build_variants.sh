#!/bin/bash
# Developer A/B builds: tools/build_variants.sh <suffix> <nvcc flags...> builds dart_env_b200/libdartb<suffix>.so with the
# fp32 topology kernels recompiled under the extra flags (the other objects are copied from the product build).
# DARTB_SO_SUFFIX=<suffix> selects the library at run time (dart_env_b200/build.py).
set -e
cd "$(dirname "$0")/.."
sfx=$1; shift
only=${VARIANT_ONLY:-walker_f,cheetah_f,hopper_f,snake_f,dartb}
rm -rf dart_env_b200/build$sfx dart_env_b200/libdartb$sfx.so
cp -r dart_env_b200/build dart_env_b200/build$sfx
DARTB_SO_SUFFIX=$sfx DARTB_NVCC_FLAGS="$*" DARTB_BUILD_ONLY=$only python -m dart_env_b200.build > /dev/null
ls -la dart_env_b200/libdartb$sfx.so
