#!/bin/bash
cd "$(dirname "$0")/../.."
for rev in f1b442f a74e101 626469c fce9ee2 a0c1910 df6b8fc; do
  echo "=== $rev"
  LIBBNR=$PWD/variants/libbnr_$rev.so timeout 120 python tools/determinism_check.py --config c3 --sweeps 60 --groups 0,0,0 2>&1 | tail -3
  LIBBNR=$PWD/variants/libbnr_$rev.so timeout 120 python tools/determinism_check.py --config c3 --chains 8 --sweeps 200 --groups 0,0,0 2>&1 | tail -3
done
echo "=== HEAD knobs"
BNR_CHOL_SERIAL=1 timeout 120 python tools/determinism_check.py --config c3 --sweeps 60 --groups 0,0,0 2>&1 | tail -2
BNR_NO_SIDE=1 timeout 120 python tools/determinism_check.py --config c3 --sweeps 60 --groups 0,0,0 2>&1 | tail -2
BNR_CHOL_A_SIDE_LO=1 timeout 120 python tools/determinism_check.py --config c3 --sweeps 60 --groups 0,0,0 2>&1 | tail -2
timeout 120 python tools/determinism_check.py --config c3 --sweeps 60 --groups 1,1,1 2>&1 | tail -2
