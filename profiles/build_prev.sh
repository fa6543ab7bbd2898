#!/bin/bash
# profiles/build_prev.sh [commit] — build an EARLIER commit's library as rala_b200/variants/librala_b200_prev.so (here, no GPU
# needed), so that the next `bench.py --ab` on the GPU box times it against the product in the same process.
set -eu
COMMIT=${1:-HEAD~1}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
TMP=$(mktemp -d)
git -C "$ROOT" worktree add -f "$TMP/src" "$COMMIT" > /dev/null
mkdir -p "$TMP/obj" "$ROOT/rala_b200/variants"
for f in $(cd "$TMP/src/rala_b200/csrc" && ls *.cu | sed 's/\.cu$//'); do
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr \
         -c "$TMP/src/rala_b200/csrc/$f.cu" -o "$TMP/obj/$f.o" &
done
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o "$ROOT/rala_b200/variants/librala_b200_prev.so" "$TMP"/obj/*.o
git -C "$ROOT" worktree remove --force "$TMP/src"
rm -rf "$TMP"
echo "rala_b200/variants/librala_b200_prev.so = $(git -C "$ROOT" rev-parse --short "$COMMIT")"
