#!/usr/bin/env bash
# One-call GPU validation of the tree, for `gpurun --timeout 1500 -- 'bash tools/gpu_validate.sh'`.
# Every stage has its own timeout and writes to gpurun_out/, so a slow box or a failing stage never
# hides the others (round 1 lost its last run to a single tight timeout).  Stages, ~10 GPU-minutes:
#   1 full parity suite (no -x, slowest tests listed)      gpurun_out/validate_tests.log
#   2 experimental kernels (B200REC_TEST_EXPERIMENTAL=1)    gpurun_out/validate_experimental.log
#   3 K6 / hash_keys micro-benchmarks, default and v2       gpurun_out/validate_dot_bench*.jsonl
#   4 smoke()                                               gpurun_out/validate_smoke.log
#   5 bench.py (N=1) and its reference arm                  gpurun_out/validate_bench*.json
#   6 ncu launch list of one bench step                     gpurun_out/validate_launches.csv
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
RC=gpurun_out/validate_rc.log
: > "$RC"
stage() {  # stage <name> <timeout-seconds> <command...>
  local name=$1 t=$2
  shift 2
  local t0=$SECONDS
  timeout "$t" "$@"
  echo "$name rc=$? seconds=$((SECONDS - t0))" >> "$RC"
}
stage tests 900 bash -c 'python -m pytest tests -m gpu -q --durations=15 > gpurun_out/validate_tests.log 2>&1'
stage experimental 240 bash -c 'B200REC_TEST_EXPERIMENTAL=1 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "v2" > gpurun_out/validate_experimental.log 2>&1'
stage dot_bench 180 bash -c 'python tools/dot_bench.py > gpurun_out/validate_dot_bench.jsonl 2> gpurun_out/validate_dot_bench.err'
stage dot_bench_v2 180 bash -c 'python tools/dot_bench.py --v2 > gpurun_out/validate_dot_bench_v2.jsonl 2> gpurun_out/validate_dot_bench_v2.err'
stage smoke 240 bash -c 'python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/validate_smoke.log 2>&1'
stage bench 420 bash -c 'python bench.py > gpurun_out/validate_bench.json 2> gpurun_out/validate_bench.err'
stage bench_ref 420 bash -c 'python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/validate_bench_ref.json 2> gpurun_out/validate_bench_ref.err'
stage launches 420 bash -c 'ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/validate_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/validate_launches.log 2>&1'
cat "$RC"
tail -n 4 gpurun_out/validate_tests.log
