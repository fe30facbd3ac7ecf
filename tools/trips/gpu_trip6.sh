#!/bin/bash
mkdir -p gpurun_out
V=build/variants
run() { AB_NAME=$1 KB2_LIB=$2 KB2_TC_CONFIG=$3 timeout 300 python tools/ab_knn.py 2>&1 | tail -1 | tee -a gpurun_out/ab.log; }
for rep in 1 2; do
  AB_NAME=v2 KB2_LIB=$PWD/$V/v2/lib/libkiez_b200.so timeout 300 python tools/ab_knn.py 2>&1 | tail -1 | tee -a gpurun_out/ab.log
  run v3 $PWD/kiez_b200/lib/libkiez_b200.so 256x32
  run v3 $PWD/kiez_b200/lib/libkiez_b200.so 256x16
  run v3_single $PWD/$V/v3_single/lib/libkiez_b200.so 256x32
  run v3_single $PWD/$V/v3_single/lib/libkiez_b200.so 256x16
  run v3_sleep $PWD/$V/v3_sleep/lib/libkiez_b200.so 256x32
  run v3_sleep $PWD/$V/v3_sleep/lib/libkiez_b200.so 256x16
  run v3_single_sleep $PWD/$V/v3_single_sleep/lib/libkiez_b200.so 256x32
  run v3_single_sleep $PWD/$V/v3_single_sleep/lib/libkiez_b200.so 256x16
done
