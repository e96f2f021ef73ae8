#!/bin/bash
# build_variant.sh NAME "FLAGS": an experimental build of libpomgpu into extpom_b200/variants/lib_NAME.so
# (kernel tuning experiments; run with POMGPU_LIB=... scripts/kprof.py)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p "$ROOT/extpom_b200/variants"
# (its own object directory: the default build's objects were compiled with other flags)
make -s -j8 -C "$ROOT/extpom_b200/csrc" "$ROOT/extpom_b200/variants/lib_$1.so" OUT="$ROOT/extpom_b200/variants/lib_$1.so" OBJ="$ROOT/extpom_b200/variants/obj_$1" LOG="$ROOT/extpom_b200/variants/ptxas_$1.log" EXTRA="$2"
