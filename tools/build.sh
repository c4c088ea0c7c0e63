#!/usr/bin/env bash
# Build libvqcpc_b200.so in-tree for sm_100a (cross-compiles without a GPU).
set -euo pipefail
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
SRC="$ROOT/vqcpc-bach_b200/csrc"
# A/B builds: VQ_EPI_KIND=k VQ_TAG=_epik tools/build.sh -> libvqcpc_b200_epik.so with its own object directory
TAG=${VQ_TAG:-}
OUT="$ROOT/vqcpc-bach_b200/libvqcpc_b200$TAG.so"
OBJ="$ROOT/build/obj$TAG"
mkdir -p "$OBJ"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -I$ROOT/include -I$SRC ${VQ_EPI_KIND:+-DVQ_EPI_KIND=$VQ_EPI_KIND}"
pids=()
objs=()
for f in "$SRC"/*.cu; do
  o="$OBJ/$(basename "${f%.cu}").o"
  objs+=("$o")
  if [[ ! -f "$o" || "$f" -nt "$o" || -n "$(find "$SRC" "$ROOT/include" \( -name '*.cuh' -o -name '*.h' \) -newer "$o" 2>/dev/null | head -1)" ]]; then
    ( $NVCC $FLAGS ${VQ_PTXAS_V:+-Xptxas -v} -c "$f" -o "$o" ) &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [[ -n "$p" ]] && wait "$p"; done
$NVCC -shared -o "$OUT" "${objs[@]}" -gencode arch=compute_100a,code=sm_100a -lcudart
echo "built $OUT"
