#!/bin/bash
# Tuning variants of the library (never a fallback: selected explicitly with BPPERM_LIB=<path>), built beside the product
# library: tools/build_variants.sh name "-DFLAG=.." [name2 "-D.."] ...  -> bulletproof-perm_b200/variants/libbpperm_<name>.so
set -e
cd "$(dirname "$0")/../bulletproof-perm_b200/csrc"
mkdir -p ../variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  make -j8 OUT=../variants/libbpperm_$name.so OBJDIR=../build_$name EXTRA="$flags" > /dev/null
  rm -rf ../build_$name
  echo built variants/libbpperm_$name.so "($flags)"
done
