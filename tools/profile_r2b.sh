#!/bin/bash
# Round-2 (second session) measurement pass on one B200: bench line, workload lines, ncu launch list of one training step,
# ncu --set full of the one-kernel coupling network (with / without side outputs) and of the dominant weight gradient.
# Every ncu capture runs after the same command has exited 0 without ncu.  Output: gpurun_out/r2b_*.
set -u
O=gpurun_out
python bench.py --steps 20 --warmup 5 > $O/r2b_bench_1gpu.json 2> $O/r2b_bench_1gpu.err || echo "bench failed"
tail -c 300 $O/r2b_bench_1gpu.json; echo
for w in rfn_J_fwd rfn_J_sample glow_cfg1 convlstm_cfg2 rfn_J_smooth_D3 rfn_D_sample; do
  timeout 200 python bench.py --workload $w --steps 20 --warmup 5 > $O/r2b_wl_$w.json 2> $O/r2b_wl_$w.err || echo "$w failed"
  python -c "import json,sys; d=json.loads(open('$O/r2b_wl_$w.json').read().strip().splitlines()[-1]); print('$w', d['value'], d['unit'], d['ms_per_step'])"
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/r2b_train_step_launches.csv \
  python bench.py --steps 2 --warmup 1 > $O/r2b_ncu_bench.log 2>&1 || echo "ncu launch list failed"
wc -l $O/r2b_train_step_launches.csv
python tools/nn_fused_bench.py one > /dev/null && timeout 120 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:coupling_nn -o $O/r2b_nn_fused_full -f python tools/nn_fused_bench.py one > $O/r2b_ncu_one.log 2>&1
python tools/nn_fused_bench.py one store > /dev/null && timeout 120 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:coupling_nn -o $O/r2b_nn_fused_store_full -f python tools/nn_fused_bench.py one store > $O/r2b_ncu_one_store.log 2>&1
python tools/wgrad_one.py 32 18 256 9 > /dev/null && timeout 120 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:wgrad -o $O/r2b_wgrad_18x256_full -f python tools/wgrad_one.py 32 18 256 9 > $O/r2b_ncu_wgrad.log 2>&1
ls -la $O/r2b_*.ncu-rep
