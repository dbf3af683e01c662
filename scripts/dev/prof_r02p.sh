set -x
python -m pytest tests -m gpu -q 2>&1 | tail -2
python bench.py > gpurun_out/r02p_bench_n1.json 2> gpurun_out/r02p_bench_n1.err; echo bench rc=$?
python bench.py --impl reference --steps 3 --warmup 1 --no-full-step > gpurun_out/r02p_ref_n1.json 2> gpurun_out/r02p_ref_n1.err; echo ref rc=$?
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02p_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02p_ncu_bench.log 2>&1; echo l1 rc=$?
SS_PROFILE_CFG=c3 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/r02p_c3_launches.csv python scripts/profile_target.py > gpurun_out/r02p_ncu6.log 2>&1; echo p6 rc=$?
