#!/bin/bash
# Build an A/B variant of the library: profiles/experiments/build_variant.sh NAME "-DFLAG=..."
#   -> profiles/experiments/variants/libb200ctl_NAME.so  (git-ignored; travels with gpurun)
# quick_time.py --lib <path> times it instead of the in-tree build.
set -e
NAME=$1; EXTRA=$2
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
SRC=$ROOT/test_isaacgym_b200/csrc
OUT=$ROOT/profiles/experiments/variants
mkdir -p "$OUT" "$SRC/build_$NAME"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden -cudart static --expt-relaxed-constexpr"
for f in runtime pd_torque servo franka franka_task peaks peer; do
  nvcc $FLAGS $EXTRA -c "$SRC/$f.cu" -o "$SRC/build_$NAME/$f.o" &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o "$OUT/libb200ctl_$NAME.so" "$SRC"/build_$NAME/*.o -ldl
rm -rf "$SRC/build_$NAME"
echo "$OUT/libb200ctl_$NAME.so"
