#!/bin/bash
# Disassembles libpm_b200.so's backplane kernels with inlining line info:
#   tools/disasm_kernel.sh out.dis
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
TMP=$(mktemp -d)
(cd "$TMP" && cuobjdump -xelf backplane_kernels "$ROOT/planetmapper_b200/libpm_b200.so" >/dev/null)
nvdisasm -gi -c "$TMP"/backplane_kernels*.cubin > "$1"
rm -rf "$TMP"
