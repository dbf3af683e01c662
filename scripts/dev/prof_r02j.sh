set -x
python bench.py > gpurun_out/r02j_bench_n1.json 2> gpurun_out/r02j_bench_n1.err; echo bench rc=$?
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02j_ref_n1.json 2> gpurun_out/r02j_ref_n1.err; echo ref rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02j_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02j_ncu_bench.log 2>&1; echo l1 rc=$?
SS_PROFILE_CFG=c3 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/r02j_c3_launches.csv python scripts/profile_target.py > gpurun_out/r02j_ncu6.log 2>&1; echo p6 rc=$?
python -m pytest tests -m gpu -q 2>&1 | tail -2
python __graft_entry__.py --smoke 2>&1 | tail -1
