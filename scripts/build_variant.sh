#!/bin/bash
# What-if build of the library under variants/<name>/ (git-ignored; objects under /tmp/vbuild/<name>): scripts/build_variant.sh name "-DFLAG1 -DFLAG2"
name=$1; flags=$2
mkdir -p /root/repo/variants/$name
PINN_BUILD_DIR=/tmp/vbuild/$name PINN_LIB_OUT=/root/repo/variants/$name/libpinn_b200.so PINN_EXTRA_NVCC_FLAGS="$flags" python pinnacle_b200/csrc/build.py > /root/repo/variants/$name/build.log 2>&1 && echo "built $name"
