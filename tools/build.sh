#!/bin/bash
# nvcc build of the product library with ptxas statistics (same flags as __graft_entry__.build)
cd /root/repo/eicos_b200/csrc || exit 1
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -Xptxas -v "$@" \
  -o ../libeicos_b200.so engine.cu capi.cu symbolic.cpp amd.cpp streams.cpp machine.cpp 2>&1 | grep -E "error|warning|Compiling|Used|spill" | sed 's/ptxas info    : //'
