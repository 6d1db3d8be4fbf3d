#!/bin/bash
cd "$(dirname "$0")/../.."
B="python bench.py --no-cpu-baseline --no-other-configs --no-strong --ess-draws 0 --steps 100"
run() { tag=$1; shift; timeout 200 "$@" > gpurun_out/r2i_$tag.json 2> gpurun_out/r2i_$tag.err; }
run c3_s4 $B --config c3
LIBBNR=$PWD/build/libbnr_s6.so run c3_s6 $B --config c3
run c5_s4 $B --config c5
LIBBNR=$PWD/build/libbnr_s6.so run c5_s6 $B --config c5
run c3x8_s4 $B --config c3 --chains 8
LIBBNR=$PWD/build/libbnr_s6.so run c3x8_s6 $B --config c3 --chains 8
LIBBNR=$PWD/build/libbnr_s6.so run c4_s6 $B --config c4
LIBBNR=$PWD/build/libbnr_s6.so run c2_s6 $B --config c2
