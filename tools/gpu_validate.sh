#!/bin/bash
# One-GPU validation pass used during round 2 (run under gpurun).  $1 = ncu|none, $2 = file prefix.
set -u
O=gpurun_out
P=${2:-r2}
mkdir -p $O
timeout 500 python -m pytest tests/test_gpu_tc_gemm.py tests/test_gpu_fullsize.py -q -rs -rf 2>&1 | grep -v "^  \|^E  \|^$" | tail -40 > $O/${P}_tests.log
timeout 900 python -m pytest tests -m gpu -q -rf --deselect tests/test_gpu_tc_gemm.py --deselect tests/test_gpu_fullsize.py 2>&1 | grep -v "^  \|^E  \|^$" | tail -25 > $O/${P}_pytest_gpu.log
timeout 200 python tools/tc_probe.py perf > $O/${P}_tc_perf.jsonl 2>&1
timeout 400 python tools/k1_tune.py warp > $O/${P}_k1_warp.jsonl 2>&1
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -2 > $O/${P}_bench_n1.json
timeout 300 python tools/dcn_bench.py --iters 5 > $O/${P}_dcn_bench.jsonl 2>&1
B200REC_TOWER=cublas timeout 300 python tools/dcn_bench.py --iters 5 >> $O/${P}_dcn_bench.jsonl 2>&1
if [ "${1:-}" = "ncu" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 300 --csv \
      --log-file $O/${P}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${P}_bench_under_ncu.log 2>&1
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 2 -c 6 \
      -o $O/${P}_prof_tc_gemm python tools/tc_probe.py perf > $O/${P}_ncu_tc.log 2>&1
fi
cat $O/${P}_tests.log; cat $O/${P}_pytest_gpu.log; cat $O/${P}_tc_perf.jsonl; cat $O/${P}_k1_warp.jsonl; cat $O/${P}_dcn_bench.jsonl | tail -3; cat $O/${P}_bench_n1.json
