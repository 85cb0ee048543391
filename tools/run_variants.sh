#!/bin/bash
# run_variants.sh PATH(call|normcounts) MB NAME...: tools/run_path.py with every build/variants/NAME.so in turn
p=$1; mb=$2; shift 2
for v in "$@"; do
  echo "variant $v"
  HIMUT_B200_LIB=$PWD/build/variants/$v.so timeout 300 python tools/run_path.py $p --contig-mb $mb --reps 4 2>&1 | tail -n 2 | cut -c1-700
done
