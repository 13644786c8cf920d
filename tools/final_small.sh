#!/bin/bash
rm -f gpurun_out/parity_trainers.txt
python -m pytest tests -m gpu -x -q -s > gpurun_out/r2c_pytest_full.log 2>&1; tail -2 gpurun_out/r2c_pytest_full.log
grep PARITY gpurun_out/r2c_pytest_full.log > gpurun_out/r2c_parity_full_size.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/r2c_bench_final.json 2> gpurun_out/r2c_bench_final.err; cut -c1-200 gpurun_out/r2c_bench_final.json
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c_launches_step_final.csv python profiles/profile_step.py > gpurun_out/r2c_ncu_launches.log 2>&1; tail -1 gpurun_out/r2c_ncu_launches.log
