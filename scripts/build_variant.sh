#!/bin/bash
# build_variant.sh NAME "FLAGS": an experimental build of libpomgpu into extpom_b200/variants/lib_NAME.so
# (kernel tuning experiments; run with POMGPU_LIB=... scripts/kprof.py)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p "$ROOT/extpom_b200/variants"
make -s -C "$ROOT/extpom_b200/csrc" OUT="$ROOT/extpom_b200/variants/lib_$1.so" LOG="$ROOT/extpom_b200/variants/ptxas_$1.log" EXTRA="$2"
