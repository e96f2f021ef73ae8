#!/bin/bash
# asan_check.sh -- the kernel bodies, the step orchestration and the halo bookkeeping under AddressSanitizer.
# Builds the host-emulated library (the same .cu sources, -DPOMGPU_EMU) with -fsanitize=address into /tmp and runs
# the emulation parity suite and the strip tests against it: every field is its own allocation, so a kernel body that
# reads or writes a row / column / level outside an array (j-1 at j=1, k+1 at k=kb, a ghost row that is not there)
# is reported.  Round 2: 0 reports over 101 tests (all step / routine / strip cases incl. the open-boundary ones).
#   asan_check.sh            AddressSanitizer
#   asan_check.sh undefined  UndefinedBehaviorSanitizer (signed overflow, static-array bounds such as the [KMAX]
#                            column vectors, misaligned access; -fno-sanitize-recover): round 2: 0 reports
set -e
SAN=${1:-address}
LIBSAN=$([ "$SAN" = address ] && echo libasan.so || echo libubsan.so)
ROOT=$(cd "$(dirname "$0")/.." && pwd)
cd "$ROOT/extpom_b200/csrc"
g++ -O1 -g -fsanitize=$SAN -fno-sanitize-recover=undefined -fno-omit-frame-pointer -ffp-contract=off -std=c++17 -fPIC -DPOMGPU_EMU -Wno-unused \
    -Wno-unknown-pragmas -shared -o /tmp/libpomgpu_emu_asan.so -x c++ pom_state.cu -x c++ pom_k_lateral.cu \
    -x c++ pom_k_external.cu -x c++ pom_k_internal.cu -x c++ pom_k_bcond.cu -x c++ pom_halo.cu -x c++ pom_forcing.cu \
    -x c++ pom_selftest.cu -x c++ pom_step.cu -lm
cd "$ROOT"
cat > /tmp/pomgpu_asan_run.py <<'PY'
import sys
sys.path.insert(0, sys.argv[1])
import tests.emu as emu
emu.build_emu = lambda: "/tmp/libpomgpu_emu_asan.so"
emu.EMU_SO = "/tmp/libpomgpu_emu_asan.so"
import pytest
sys.exit(pytest.main(["tests/test_strips.py", "tests/test_emu_parity.py", "-x", "-q", "-m", "not gpu",
                      "-k", "not gloo and not two_processes", "-p", "no:cacheprovider"]))
PY
LD_PRELOAD=$(gcc -print-file-name=$LIBSAN) ASAN_OPTIONS=detect_leaks=0:halt_on_error=0 UBSAN_OPTIONS=print_stacktrace=1 \
    python /tmp/pomgpu_asan_run.py "$ROOT" 2>&1 | tee /tmp/pomgpu_asan.log | tail -3
echo "sanitizer reports: $(grep -c 'ERROR: AddressSanitizer\|runtime error' /tmp/pomgpu_asan.log || true)"
