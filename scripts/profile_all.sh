# Round-2 ncu captures (run under gpurun, one GPU).  Outputs in gpurun_out/; summaries are made here with scripts/summarize_ncu.py.
set -x
# launch lists (every launch with its device time; cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 120 --csv --log-file gpurun_out/r2_launches_quad.csv python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/r2_bench_under_ncu_quad.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 120 --csv --log-file gpurun_out/r2_launches_maze.csv python bench.py --workload maze3d --steps 20 --warmup 5 --no-extras > gpurun_out/r2_bench_under_ncu_maze.log 2>&1
# full captures
ncu --set full --clock-control none --import-source on -k regex:quad_step_wide -s 12 -c 1 -o gpurun_out/prof_r2_quad_step_wide_65k python scripts/profile_quad.py 65536 20 > gpurun_out/ncu_r2_q1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:quad_stream -s 6 -c 1 -o gpurun_out/prof_r2_quad_stream_4m python scripts/profile_quad.py 4194304 10 > gpurun_out/ncu_r2_q2.log 2>&1
MGB_PACKED=1 ncu --set full --clock-control none --import-source on -k regex:quad_step2 -s 12 -c 1 -o gpurun_out/prof_r2_quad_step2_packed_65k python scripts/profile_quad.py 65536 20 > gpurun_out/ncu_r2_q3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:maze3d_step_kernel -s 150 -c 1 -o gpurun_out/prof_r2_maze3d_step_1024 python scripts/profile_maze.py 1024 200 > gpurun_out/ncu_r2_m1.log 2>&1
MGB_MAZE_CACHE=0 ncu --set full --clock-control none --import-source on -k regex:maze3d_kernel -s 4 -c 1 -o gpurun_out/prof_r2_maze3d_direct_1024 python scripts/profile_maze.py 1024 8 > gpurun_out/ncu_r2_m2.log 2>&1
for f in gpurun_out/ncu_r2_*.log; do tail -n 2 "$f"; done
