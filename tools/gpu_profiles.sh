mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -1 gpurun_out/bench_n1.json | cut -c1-330
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_n1.json 2>&1
timeout 900 python tests/gpu_sweep.py 1048576 > gpurun_out/sweep.log 2>&1
timeout 600 python tests/gpu_sweep.py 262144 p521 > gpurun_out/sweep_p521.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/bench_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k256_verify_kernel -s 2 -c 1 -o gpurun_out/verify_full python tests/gpu_quick.py 1048576 > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:ed25519_verify_kernel -s 1 -c 1 -o gpurun_out/ed25519_full python tests/gpu_sweep.py 262144 ed25519 > gpurun_out/ncu_ed.log 2>&1
du -sh gpurun_out; ls -la gpurun_out | awk '{print $5, $9}' | tr '\n' ' '
