#!/bin/bash
# Builds lib/libpgmvae.so for sm_100a (cross-compiles without a GPU).
set -e
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
mkdir -p "$HERE/lib" "$HERE/build"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall ${PGMVAE_NVCC_FLAGS}"
objs=""
pids=""
for f in "$HERE"/csrc/*.cu; do
  o="$HERE/build/$(basename "${f%.cu}").o"
  objs="$objs $o"
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ -n "$(find "$HERE/csrc" "$HERE/../include" \( -name '*.cuh' -o -name '*.h' \) -newer "$o")" ]; then
    $NVCC $FLAGS -c "$f" -o "$o" &
    pids="$pids $!"
  fi
done
for p in $pids; do wait $p; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o "$HERE/lib/libpgmvae.so" $objs -cudart static -ldl -lpthread -lrt
echo "built $HERE/lib/libpgmvae.so"
