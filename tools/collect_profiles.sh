#!/bin/bash
# Round-2 evidence run (one GPU): launch list of the bench command and one `ncu --set full` capture per kernel.
# Everything lands in gpurun_out/; tools/summarize_ncu.py turns the reports into the text files under profiles/.
cd "$(dirname "$0")/.."
O=gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > $O/r2_plain_bench.json 2> $O/r2_plain_bench.err || exit 1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file $O/r2_launches_cfg2.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > $O/ncu_launches.log 2>&1
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:merge_stream_kernel -s 2 -c 1 -o $O/r2_merge_stream python tools/run_merge.py 0.05 3 1 1 > $O/ncu_a0.log 2>&1
$NCU -k regex:merge_staged_kernel -s 2 -c 1 -o $O/r2_merge_staged python tools/run_merge.py 0.05 3 1 1 - 2 > $O/ncu_a.log 2>&1
$NCU -k regex:merge_stream_lut -s 2 -c 1 -o $O/r2_merge_stream_lut python tools/run_merge.py 0.05 3 1 1 lut > $O/ncu_b0.log 2>&1
$NCU -k regex:merge_staged_lut -s 2 -c 1 -o $O/r2_merge_staged_lut python tools/run_merge.py 0.05 3 1 1 lut 2 > $O/ncu_b.log 2>&1
$NCU -k regex:dark_scan -s 2 -c 1 -o $O/r2_dark_scan python tools/run_merge.py 0.05 3 1 1 > $O/ncu_c.log 2>&1
$NCU -k regex:roi_partial -s 1 -c 1 -o $O/r2_roi python tools/run_merge.py 0.05 3 1 1 > $O/ncu_d.log 2>&1
$NCU -k regex:energy_partial -s 3 -c 1 -o $O/r2_k4_nostd python tools/run_k4.py 0 3 > $O/ncu_e.log 2>&1
$NCU -k regex:energy_partial -s 3 -c 1 -o $O/r2_k4_std python tools/run_k4.py 1 3 > $O/ncu_f.log 2>&1
$NCU -k regex:energy_tail -s 3 -c 1 -o $O/r2_k4_tail python tools/run_k4.py 0 3 > $O/ncu_g.log 2>&1
$NCU -k regex:welford_stack_lut -s 1 -c 1 -o $O/r2_k3_lut python tools/run_k3.py 1 2 > $O/ncu_h.log 2>&1
$NCU -k regex:welford_stack_u8 -s 1 -c 1 -o $O/r2_k3_u8 python tools/run_k3.py 0 2 > $O/ncu_i.log 2>&1
$NCU -k regex:pair_stats -s 1 -c 1 -o $O/r2_pair python tools/run_pair_stats.py > $O/ncu_j.log 2>&1
$NCU -k regex:linearize -c 1 -o $O/r2_k1 python -c "
import torch, sys
sys.path.insert(0, '.')
import bench
from camera_linearity_b200 import ops
g = torch.Generator(device='cuda').manual_seed(1)
icrf, diff = bench.icrf_tables(3)
icrf, diff = torch.from_numpy(icrf).cuda(), torch.from_numpy(diff).cuda()
dn = torch.randint(0, 256, (2160, 3840, 3), generator=g, device='cuda', dtype=torch.uint8)
sd = torch.rand((2160, 3840, 3), generator=g, device='cuda', dtype=torch.float64) * 0.02
for _ in range(3): ops.linearize(dn, sd, icrf, diff)
torch.cuda.synchronize()" > $O/ncu_k.log 2>&1
ls -la $O/*.ncu-rep | tail -20
