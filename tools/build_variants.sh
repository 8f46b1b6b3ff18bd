#!/bin/bash
# tuning builds of libsodm.so: name:flags pairs -> .variants/libsodm_<name>.so (travels to the GPU box; load with PYSODM_B200_LIB)
# VARIANT_SRC=<file stem> picks the source the flags apply to (default sodm_jump)
set -e
cd "$(dirname "$0")/.."
mkdir -p .variants
for v in "$@"; do
  name="${v%%:*}"; flags="${v#*:}"
  ( mkdir -p .variants/obj_$name
    for f in pysodm_b200/csrc/*.cu; do
      b=$(basename $f .cu)
      if [ "$b" = "${VARIANT_SRC:-sodm_jump}" ] || [ ! -f pysodm_b200/csrc/_obj/$b.o ]; then
        nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC $flags -c $f -o .variants/obj_$name/$b.o >> .variants/build_$name.log 2>&1 || echo "FAILED $name $b"
      else
        cp pysodm_b200/csrc/_obj/$b.o .variants/obj_$name/$b.o
      fi
    done
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -o .variants/libsodm_$name.so .variants/obj_$name/*.o -cudart static >> .variants/build_$name.log 2>&1 || echo "FAILED link $name"
    rm -rf .variants/obj_$name ) &
done
wait
ls -la .variants/*.so
