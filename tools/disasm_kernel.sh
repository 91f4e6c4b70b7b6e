#!/bin/bash
# Disassembles one translation unit of libpm_b200.so with inlining line info:
#   tools/disasm_kernel.sh out.dis [backplane_kernels|gather_kernels|proj_kernels]
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
UNIT="${2:-backplane_kernels}"
TMP=$(mktemp -d)
(cd "$TMP" && cuobjdump -xelf "$UNIT" "$ROOT/planetmapper_b200/libpm_b200.so" >/dev/null)
nvdisasm -gi -c "$TMP"/"$UNIT"*.cubin > "$1"
rm -rf "$TMP"
