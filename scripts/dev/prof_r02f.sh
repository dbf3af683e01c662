set -x
python bench.py > gpurun_out/r02f_bench_n1.json 2> gpurun_out/r02f_bench_n1.err; echo bench rc=$?
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02f_ref_n1.json 2> gpurun_out/r02f_ref_n1.err; echo ref rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02f_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02f_ncu_bench.log 2>&1; echo l1 rc=$?
ncu --set full --import-source on --clock-control none -k regex:mpc_rollout_tc_kernel -c 1 -f -o gpurun_out/r02f_prof_tc python scripts/profile_target.py > gpurun_out/r02f_ncu1.log 2>&1; echo p1 rc=$?
ncu --set full --import-source on --clock-control none -k regex:kde_pairs_tc -c 1 -f -o gpurun_out/r02f_prof_kde python scripts/profile_target.py > gpurun_out/r02f_ncu2.log 2>&1; echo p2 rc=$?
ncu --set full --import-source on --clock-control none -k regex:mpc_tail -c 1 -f -o gpurun_out/r02f_prof_tail python scripts/profile_target.py > gpurun_out/r02f_ncu3.log 2>&1; echo p3 rc=$?
SS_PROFILE_CFG=c3 ncu --set full --import-source on --clock-control none -k regex:quad -s 2 -c 1 -f -o gpurun_out/r02f_prof_quad python scripts/profile_target.py > gpurun_out/r02f_ncu4.log 2>&1; echo p4 rc=$?
SS_PROFILE_CFG=mt ncu --set full --import-source on --clock-control none -k regex:mt19937 -c 8 -f -o gpurun_out/r02f_prof_mt python scripts/profile_target.py > gpurun_out/r02f_ncu5.log 2>&1; echo p5 rc=$?
SS_PROFILE_CFG=c3 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/r02f_c3_launches.csv python scripts/profile_target.py > gpurun_out/r02f_ncu6.log 2>&1; echo p6 rc=$?
