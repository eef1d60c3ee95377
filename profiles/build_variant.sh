#!/bin/bash
# Build libcfm_b200.so of another git ref (or the working tree with extra -D flags) into profiles/ab/<name>.so for
# same-box A/B runs: CFM_B200_LIB=profiles/ab/<name>.so python profiles/quick_perf.py
#   usage: profiles/build_variant.sh <name> <git-ref|WORK> [extra nvcc flags]
set -e
name=$1; ref=$2; shift 2
root=$(cd "$(dirname "$0")/.." && pwd)
pkg=image-inpainting-and-super-resolution-using-diffusion-models-and-conditional-flow-matching_b200
tmp=$(mktemp -d)
if [ "$ref" = WORK ]; then cp -r "$root/$pkg" "$tmp/$pkg"; mkdir -p "$tmp/include"; cp "$root"/include/*.h "$tmp/include/";
else (cd "$root" && git archive "$ref" "$pkg/csrc" include) | tar -x -C "$tmp"; fi
mkdir -p "$root/profiles/ab" "$tmp/obj"
for f in "$tmp/$pkg"/csrc/*.cu; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -c "$f" -o "$tmp/obj/$(basename "$f" .cu).o" &
done
wait
/usr/local/cuda/bin/nvcc -shared -o "$root/profiles/ab/$name.so" "$tmp"/obj/*.o -lcudart
rm -rf "$tmp"
echo "built profiles/ab/$name.so"
