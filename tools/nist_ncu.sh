export EB200_CACHE=/tmp/eb200_cache
summ() {
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$1 -s $5 -c 1 -o /tmp/$3 python bench.py --steps 1 --only $2 > gpurun_out/ncu_$3.log 2>&1
  python tools/ncu_summary.py /tmp/$3.ncu-rep gpurun_out/r02_ncu_$3_summary.txt "ncu --set full --clock-control none, $4 (python bench.py --steps 1 --only $2), round 2" > /dev/null 2>&1
}
summ sw_verify_kernel p256_verify p256_verify "sw_verify_kernel<P256>, N = 2^20" 3
summ sw_verify_kernel p384_verify p384_verify "sw_verify_kernel<P384>, N = 2^20" 3
summ sw_verify_kernel p521_verify p521_verify "sw_verify_kernel<P521>, N = 2^18" 3
