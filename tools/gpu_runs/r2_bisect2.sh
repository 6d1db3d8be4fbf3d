#!/bin/bash
cd "$(dirname "$0")/../.."
for v in V1 V2; do
  echo "=== $v"
  LIBBNR=$PWD/variants/libbnr_$v.so timeout 120 python tools/determinism_check.py --config c3 --sweeps 60 --groups 0,0,0 2>&1 | tail -2
  LIBBNR=$PWD/variants/libbnr_$v.so timeout 120 python tools/determinism_check.py --config c3 --chains 8 --sweeps 200 --groups 0,0,0 2>&1 | tail -2
done
