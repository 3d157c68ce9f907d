set -x
mkdir -p gpurun_out
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo smoke=$?
python bench.py > gpurun_out/bench_r1_n1.json 2> gpurun_out/bench_r1_n1.err; echo bench=$?
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_r1_ref.json 2> gpurun_out/bench_r1_ref.err; echo ref=$?
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_search|k_residual" -c 16 --csv --log-file gpurun_out/launches_cfg2.csv python scripts/profile_once.py velodyne_30k_1m 2 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_search -s 2 -c 1 -f -o gpurun_out/prof_search_cfg2 python scripts/profile_once.py velodyne_30k_1m 2 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_residual -s 5 -c 1 -f -o gpurun_out/prof_residual_cfg2 python scripts/profile_once.py velodyne_30k_1m 2 1 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_undistort|k_vg_" -c 24 --csv --log-file gpurun_out/launches_frontend.csv python scripts/frontend_bench.py 120000 2 0 > /dev/null 2>&1
tail -1 gpurun_out/bench_r1_n1.json | cut -c1-1500
tail -1 gpurun_out/bench_r1_ref.json | cut -c1-600
tail -3 gpurun_out/smoke.log
