#!/bin/bash
# A/B builds of the library with different -D switches -> variants/<name>.so (git-ignored; they travel with gpurun)
set -e
cd "$(dirname "$0")/.."
mkdir -p variants
rm -f variants/*.so
build() {  # name, flags...
  name=$1; shift
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC "$@" \
       -o variants/$name.so midastouch_b200/csrc/midas_b200.cu &
}
while read -r name flags; do
  [ -z "$name" ] && continue
  build $name $flags
done <<< "${VARIANT_SPEC}"
wait
ls -la variants/
