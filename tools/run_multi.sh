export WRNN_SPIN_DEADLINE_MS=5000
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2g_smoke.txt 2>&1; echo "smoke rc=$?" >> gpurun_out/r2g_smoke.txt; tail -12 gpurun_out/r2g_smoke.txt
timeout 300 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/r2g_gputests_multi.txt 2>&1; tail -3 gpurun_out/r2g_gputests_multi.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2g_bench_cfg3ref_2gpu.json 2> gpurun_out/r2g_bench_2gpu.err; tail -c 400 gpurun_out/r2g_bench_cfg3ref_2gpu.json
