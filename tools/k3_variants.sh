#!/bin/bash
# builds alternative libsdrb200.so files with other settings of the k2a_v3 tuning switches (kernels_v3.cuh) into
# sdrreceiver_b200/variants/ -- loaded with SDRB_LIB=... by bench.py / tools/variant_sweep.sh
set -e
cd "$(dirname "$0")/../sdrreceiver_b200/csrc"
mkdir -p ../variants
NV="/usr/local/cuda/bin/nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC"
for v in ${K3_VARIANTS:-"nopipe:-DK3_PIPE=0" "unroll1:-DK3_VFO_UNROLL=1" "unroll2:-DK3_VFO_UNROLL=2" "scalarcmul:-DK3_PACKED_CMUL=0"}; do
  name=${v%%:*}; flags=${v#*:}
  $NV $flags -c -o /tmp/api_$name.o api.cu 2>/dev/null
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o ../variants/lib_$name.so /tmp/api_$name.o plan_host.o publisher.o ingest.o facade.o -ldl -lpthread
  echo built $name
done
