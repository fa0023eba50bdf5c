export WRNN_SPIN_DEADLINE_MS=5000
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2g_gputests.txt 2>&1; tail -3 gpurun_out/r2g_gputests.txt
timeout 400 python bench.py > gpurun_out/r2g_bench_cfg3ref_default.json 2> gpurun_out/r2g_bench_default.err; tail -c 600 gpurun_out/r2g_bench_cfg3ref_default.json
timeout 200 python bench.py --workload cfg1 > gpurun_out/r2g_bench_cfg1.json 2>/dev/null
timeout 200 python bench.py --workload cfg1x60 > gpurun_out/r2g_bench_cfg1x60.json 2>/dev/null
WRNN_RS_CALIBRATE=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2g_launches_cfg3ref.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/ncu_l5.log 2>&1
WRNN_RS_CALIBRATE=0 timeout 400 ncu --set full --clock-control none --import-source on -k regex:wrnn_loop_rs -c 1 -o gpurun_out/r2g_loop_rs_cfg3ref python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-extras > gpurun_out/ncu_f5.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2g_smoke.txt 2>&1; echo "smoke rc=$?" >> gpurun_out/r2g_smoke.txt
