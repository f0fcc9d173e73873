#!/bin/bash
# One GPU call that refreshes the evidence kept under profiles/ (run through gpurun; outputs land in gpurun_out/$TAG_*).
#   tools/capture_profiles.sh TAG
# 1. GPU test suite   2. bench.py (plain, never under a profiler)   3. ncu launch list of a render-only bench step and of
# training steps   4. ncu --set full: the fine-pass network query of the bench, the training kernels, every stage kernel
TAG=${1:-cap}
O=gpurun_out
NCU_FULL="ncu --set full --clock-control none --import-source on"
NCU_LIST="ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none --csv"
python -m pytest tests -q -m gpu > $O/${TAG}_pytest.log 2>&1; tail -3 $O/${TAG}_pytest.log
python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err || tail -5 $O/${TAG}_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_reference.json 2>> $O/${TAG}_bench.err
$NCU_LIST -c 400 --log-file $O/${TAG}_render_launches.csv python bench.py --only-render --no-cpu-baseline --steps 2 --warmup 3 > $O/${TAG}_ncu_render.log 2>&1
$NCU_LIST -c 600 --log-file $O/${TAG}_train_launches.csv python tools/bench_train.py --steps 3 --warmup 2 > $O/${TAG}_ncu_train.log 2>&1
# fine pass = every second k_mlp_tc launch; skip the warm-up ones (3 warm-up steps x 2 launches + the coarse launch of the timed step)
$NCU_FULL -k regex:k_mlp_tc -s 7 -c 1 -o $O/${TAG}_fine python bench.py --only-render --no-cpu-baseline --steps 1 --warmup 3 > $O/${TAG}_ncu_fine.log 2>&1
$NCU_FULL -k regex:'k_mlp_tc|k_mlp_bwd_pipe' -s 8 -c 4 -o $O/${TAG}_train python tools/bench_train.py --steps 2 --warmup 2 > $O/${TAG}_ncu_trainfull.log 2>&1
$NCU_FULL -k regex:'k_composite|k_importance|k_stratified' -o $O/${TAG}_stages python tools/stage_ncu.py > $O/${TAG}_ncu_stages.log 2>&1
ls -la $O/${TAG}_*
