#!/bin/bash
# Final evidence run of round 2 (one GPU): GPU test suite, the default bench line, the launch list of the bench
# command and one `ncu --set full` capture for each kernel that changed in the last session.
cd "$(dirname "$0")/.."
O=gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -4 > $O/r2h_pytest.log; cat $O/r2h_pytest.log
python bench.py > $O/r2h_bench.json 2> $O/r2h_bench.err || { tail -5 $O/r2h_bench.err; exit 1; }
python tools/show_bench.py $O/r2h_bench.json 2>/dev/null | head -12
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file $O/r2h_launches_cfg2.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > $O/ncu_launches.log 2>&1
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:merge_stream_kernel -s 2 -c 1 -o $O/r2h_merge_stream python tools/run_merge.py 0.05 3 1 1 > $O/ncu_a0.log 2>&1
$NCU -k regex:merge_stream_lut -s 2 -c 1 -o $O/r2h_merge_stream_lut python tools/run_merge.py 0.05 3 1 1 lut > $O/ncu_b0.log 2>&1
$NCU -k regex:energy_partial -s 3 -c 1 -o $O/r2h_k4_std python tools/run_k4.py 1 3 > $O/ncu_f.log 2>&1
$NCU -k regex:pair_stats -s 1 -c 1 -o $O/r2h_pair python tools/run_pair_stats.py > $O/ncu_j.log 2>&1
$NCU -k regex:merge_wide_pipe -s 1 -c 1 -o $O/r2h_wide_pipe python tools/run_cfg5.py 0 2 > $O/ncu_w.log 2>&1
ls -la $O/r2h_*.ncu-rep
