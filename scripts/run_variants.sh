#!/bin/bash
# run kprof for the default library and every variant; usage: run_variants.sh "kernel,names" [n kb]
cd "$(dirname "$0")/.."
export KPROF_ONLY="$1"
N=${2:-1024}; KB=${3:-41}
python scripts/kprof.py $N $KB
for f in extpom_b200/variants/lib_*.so; do POMGPU_LIB=$PWD/$f python scripts/kprof.py $N $KB; done
