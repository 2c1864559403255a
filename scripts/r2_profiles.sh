#!/bin/bash
# Round-2 evidence run (GPU box, one GPU): launch list of the bench command, full ncu captures of the chain kernel on
# a loader-group launch (8 views), of the mix kernel and of the OA-Loss similarity kernels, in-kernel accounting.
set -x
O=gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/r2_launches_bench.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/r2_bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:oamix_chain -s 2 -c 1 -f -o $O/r2_chain \
  python scripts/profile_path.py oamix 3 8 > $O/r2_chain.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mix_kernel -s 2 -c 1 -f -o $O/r2_mix \
  python scripts/profile_path.py oamix 3 8 > $O/r2_mix.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sim_ -s 2 -c 2 -f -o $O/r2_loss \
  python scripts/profile_path.py loss 2 > $O/r2_loss.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 14 -c 40 --csv --log-file $O/r2_launches_loss.csv \
  python scripts/profile_path.py loss 3 > /dev/null 2>&1
timeout 200 python scripts/chain_stats.py 6 8 > $O/r2_cta_time_by_kind.txt 2>&1
timeout 200 python scripts/trace_items.py 8 > $O/r2_critical_path.txt 2>&1
ls -la $O | tail -12
