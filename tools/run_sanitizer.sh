#!/bin/bash
# compute-sanitizer over the mbarrier / TMEM / TMA kernels at small shapes; logs -> gpurun_out/r2_sanitizer_<tool>_<case>.log
# (racecheck only where the kernel synchronises through CTA barriers; it does not model mbarrier / async-proxy ordering)
run() {
  tool=$1; c=$2
  log=gpurun_out/r2_sanitizer_${tool}_${c}.log
  timeout 90 compute-sanitizer --tool $tool --print-limit 10 python tools/sanitize.py $c > $log 2>&1
  echo "rc=$? tool=$tool case=$c" >> $log
  tail -3 $log | tr '\n' ' '; echo
}
for c in wide2 sweep_umma predict_umma traj_narrow train_umma64 train_umma128; do run memcheck $c; done
for c in traj_narrow wide2 predict_umma; do run racecheck $c; done
