mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
for gw in 20 22 24; do echo "== GW $gw"; EB200_LIB=build_variants/lib_gw$gw.so timeout 300 python tests/gpu_quick.py 1048576 2>&1 | tail -2; done > gpurun_out/gw_variants.log 2>&1
grep -E "==|main_kernel" gpurun_out/gw_variants.log | sed -e 's/"wall_ms.*//'
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -1 gpurun_out/bench_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/bench_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k256_verify_kernel -s 1 -c 1 -o gpurun_out/verify_full python tests/gpu_quick.py 262144 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -12
