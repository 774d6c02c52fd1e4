#!/bin/bash
# One-GPU validation pass used during round 2 (run under gpurun): kernel tests, the oracle-backed
# headline-config tests, the bench line, the per-launch list and one full ncu capture of the
# tcgen05 GEMM.  Outputs land in gpurun_out/.
set -u
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_tc_gemm.py tests/test_gpu_fullsize.py -q -rs 2>&1 | tail -15 > $O/r2b_tests.log
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -2 > $O/r2b_bench_n1.json
if [ "${1:-}" = "ncu" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 600 --csv \
      --log-file $O/r2b_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/r2b_bench_under_ncu.log 2>&1
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 20 -c 6 \
      -o $O/r2b_prof_tc_gemm python tools/tc_probe.py perf > $O/r2b_ncu_tc.log 2>&1
fi
cat $O/r2b_tests.log; cat $O/r2b_bench_n1.json
