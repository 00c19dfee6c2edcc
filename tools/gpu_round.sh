#!/bin/bash
# One GPU-box visit: tests, bench, launch list, ncu --set full of the kernels named in $NCU_KERNELS.
# Usage (under gpurun): bash tools/gpu_round.sh [tests] [bench] [launches] [ncu]
set -u
mkdir -p gpurun_out
for what in "$@"; do
case $what in
tests)
  timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
  tail -5 gpurun_out/pytest_gpu.log; grep -E "^\[(parity|stage|streaming)" gpurun_out/pytest_gpu.log | tail -80 ;;
bench)
  timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench rc=$?"; tail -c 6000 gpurun_out/bench_c2.json; tail -5 gpurun_out/bench_c2.err ;;
benchref)
  timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.json ;;
launches)
  timeout 600 python tools/profile_step.py --warmup 1 --steps 1 > gpurun_out/plain.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
      python tools/profile_step.py --warmup 1 --steps 1 > gpurun_out/ncu.log 2>&1; echo "launch list rc=$?"; tail -3 gpurun_out/plain.log ;;
ncu)
  timeout 600 python tools/profile_step.py --warmup 1 --steps 1 > gpurun_out/plain2.log 2>&1 &&
  timeout 1500 ncu --set full --clock-control none --import-source on -k "regex:${NCU_KERNELS:-istft|groupnorm|rownorm|fsq_im2col}" \
      -s ${NCU_SKIP:-0} -c ${NCU_COUNT:-6} -f -o gpurun_out/${NCU_OUT:-prof} python tools/profile_step.py --warmup 1 --steps 1 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full.log ;;
esac
done
