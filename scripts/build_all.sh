#!/bin/bash
# Builds every native artefact (product library, stress/prof variants, host test lib, oracle) from the repo root.
set -e
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build()" 2>&1 | grep -E " error|Error|Traceback" && exit 1
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --shared -Xcompiler -fPIC,-pthread"
if [ "$1" = "prof" ]; then $NV -DD4_PROF -o deft4j_b200/libdeft4cu_prof.so deft4j_b200/csrc/deft4cu.cu deft4j_b200/csrc/png_front.cpp deft4j_b200/csrc/zip_front.cpp deft4j_b200/csrc/gz_front.cpp 2>&1 | grep -E " error" || true; fi
ls -la --time-style=+%T deft4j_b200/*.so | awk '{print $6, $7}'
