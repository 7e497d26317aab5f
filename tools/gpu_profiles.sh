# Final profile pass: bench lines, sweep, launch list, and ncu summaries condensed on the box (the .ncu-rep files
# are deleted there: gpurun only brings back 64 MiB).
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -1 gpurun_out/bench_n1.json | cut -c1-330
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_n1.json 2>&1
timeout 900 python tests/gpu_sweep.py 1048576 > gpurun_out/sweep.log 2>&1
timeout 600 python tests/gpu_sweep.py 262144 p521 > gpurun_out/sweep_p521.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/bench_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k256_verify_kernel -s 2 -c 1 -o /tmp/verify_full python tests/gpu_quick.py 1048576 > gpurun_out/ncu_full.log 2>&1
python tools/ncu_summary.py /tmp/verify_full.ncu-rep gpurun_out/ncu_verify_summary.txt "ncu --set full --clock-control none, k256_verify_kernel, one 2^18-signature chunk of tests/gpu_quick.py 1048576, round 1 final kernel" > /dev/null 2>&1
ncu -i /tmp/verify_full.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h,v=rows[0],rows[2]
d=dict(zip(h,v)); print({k:d[k] for k in ('dram__bytes_read.sum','dram__bytes_write.sum','gpu__time_duration.sum') if k in d})" > gpurun_out/ncu_verify_traffic.txt
timeout 600 ncu --set full --clock-control none -k regex:ed25519_verify_kernel -s 1 -c 1 -o /tmp/ed_full python tests/gpu_sweep.py 262144 ed25519 > gpurun_out/ncu_ed.log 2>&1
python tools/ncu_summary.py /tmp/ed_full.ncu-rep gpurun_out/ncu_ed25519_summary.txt "ncu --set full --clock-control none, ed25519_verify_kernel, N = 2^18 (tests/gpu_sweep.py 262144 ed25519), round 1 final kernel" > /dev/null 2>&1
du -sh gpurun_out; ls gpurun_out | tr '\n' ' '
