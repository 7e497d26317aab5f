# One gpurun call used at the end of a work session: parity, bench, sweep, ncu evidence.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -1 gpurun_out/bench_n1.json | cut -c1-700
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_n1.json 2>&1; tail -1 gpurun_out/bench_ref_n1.json | cut -c1-200
timeout 1200 python tests/gpu_sweep.py 1048576 secp256k1,p256,p384,ed25519,ed25519_msgs,curve25519,k256_sign,k256_recover,k256_mul,k256_mul_add,k256_mul_g,p256_sign,p384_sign > gpurun_out/sweep.log 2>&1; cut -c1-230 gpurun_out/sweep.log | sed -e 's/"h2d_ms[^m]*"main/"main/' | tail -14
timeout 600 python tests/gpu_sweep.py 262144 p521 > gpurun_out/sweep_p521.log 2>&1; cut -c1-230 gpurun_out/sweep_p521.log | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/bench_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k256_verify_kernel -s 2 -c 1 -o gpurun_out/verify_full python tests/gpu_quick.py 1048576 > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ed25519_verify_kernel -s 1 -c 1 -o gpurun_out/ed25519_full python tests/gpu_sweep.py 262144 ed25519 > gpurun_out/ncu_ed.log 2>&1
ls gpurun_out | tr '\n' ' '
