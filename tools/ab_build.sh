#!/bin/bash
# tools/ab_build.sh <tag> <extra nvcc flags...>: build a variant of the library into polyphemus_b200/lib_<tag>/ for A/B runs
tag=$1; shift
PB_EXTRA_NVCC="$*" python - <<PY
import os, shutil
from polyphemus_b200 import build
build.build(force=True)
os.makedirs("gpurun_ab", exist_ok=True)
shutil.copy(build.LIB_PATH, "gpurun_ab/lib_$tag.so")
PY
