# One gpurun call used at the end of a work session: parity, variants, sweep, ncu evidence.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
for v in b64 gw20 gw20_b64; do [ -f build_variants/lib_$v.so ] || continue; echo "== $v"; EB200_LIB=build_variants/lib_$v.so timeout 300 python tests/gpu_quick.py 1048576 2>&1 | tail -2; done > gpurun_out/variants.log 2>&1
grep -E "==|main_kernel" gpurun_out/variants.log | sed -e 's/"wall_ms.*//' -e 's/"h2d_ms.*"main/"main/'
timeout 900 python tests/gpu_sweep.py 1048576 > gpurun_out/sweep.log 2>&1; cut -c1-230 gpurun_out/sweep.log | sed -e 's/"h2d_ms[^m]*"main/"main/' | tail -12
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k256_verify_kernel -s 2 -c 1 -o gpurun_out/verify_full python tests/gpu_quick.py 1048576 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -8
