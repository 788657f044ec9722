#!/usr/bin/env bash
# Builds ecamp_b200/lib/libecamp_b200.so for sm_100a (cross-compiles without a GPU).
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
out="$here/../lib"
mkdir -p "$out" "$here/build"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden)
objs=()
pids=()
for src in "$here"/*.cu; do
  obj="$here/build/$(basename "${src%.cu}").o"
  objs+=("$obj")
  if [[ ! -f "$obj" || "$src" -nt "$obj" || -n "$(find "$here" -maxdepth 1 \( -name '*.cuh' -o -name '*.h' \) -newer "$obj" 2>/dev/null)" || "$here/../../include/ecamp_b200.h" -nt "$obj" ]]; then
    "$NVCC" "${FLAGS[@]}" -c "$src" -o "$obj" &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [[ -n "$p" ]] && wait "$p"; done
"$NVCC" -shared -o "$out/libecamp_b200.so" "${objs[@]}" -lcudart
echo "built $out/libecamp_b200.so"
