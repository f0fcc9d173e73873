O=gpurun_out
python -m pytest tests -q -m gpu > $O/r2c_pytest.log 2>&1; tail -2 $O/r2c_pytest.log
python bench.py > $O/r2c_bench.json 2> $O/r2c_bench.err || tail -5 $O/r2c_bench.err
ncu --set full --clock-control none --import-source on -k regex:'k_composite|k_importance|k_stratified' -o $O/r2c_stages python tools/stage_ncu.py > $O/r2c_ncu_stages.log 2>&1
ls -la $O/r2c_*
