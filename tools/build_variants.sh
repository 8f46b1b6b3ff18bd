#!/bin/bash
# tuning builds of libsodm.so: name:flags pairs
set -e
cd "$(dirname "$0")/.."
mkdir -p .variants
rm -f .variants/*.so
for v in "$@"; do
  name="${v%%:*}"; flags="${v#*:}"
  ( nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC $flags -shared \
      -o .variants/libsodm_$name.so pysodm_b200/csrc/*.cu -cudart static > .variants/build_$name.log 2>&1 || echo "FAILED $name" ) &
done
wait
ls .variants/*.so
