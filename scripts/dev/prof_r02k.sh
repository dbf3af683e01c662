ncu --set full --import-source on --clock-control none -k regex:mpc_rollout_tc_kernel -c 1 -f -o gpurun_out/r02k_prof_tc python scripts/profile_target.py > gpurun_out/r02k_ncu1.log 2>&1; echo p1 rc=$?
ncu --set full --import-source on --clock-control none -k regex:kde_pairs_tc -c 1 -f -o gpurun_out/r02k_prof_kde python scripts/profile_target.py > gpurun_out/r02k_ncu2.log 2>&1; echo p2 rc=$?
ncu --set full --import-source on --clock-control none -k regex:mpc_tail -c 1 -f -o gpurun_out/r02k_prof_tail python scripts/profile_target.py > gpurun_out/r02k_ncu3.log 2>&1; echo p3 rc=$?
SS_PROFILE_CFG=c3 ncu --set full --import-source on --clock-control none -k regex:quad -s 2 -c 1 -f -o gpurun_out/r02k_prof_quad python scripts/profile_target.py > gpurun_out/r02k_ncu4.log 2>&1; echo p4 rc=$?
SS_PROFILE_CFG=mt ncu --set full --import-source on --clock-control none -k regex:mt19937 -c 8 -f -o gpurun_out/r02k_prof_mt python scripts/profile_target.py > gpurun_out/r02k_ncu5.log 2>&1; echo p5 rc=$?
