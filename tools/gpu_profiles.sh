# Round-2 profile pass (run under gpurun, one GPU): the launch list of the bench command, one `ncu --set full`
# capture per dominant kernel condensed on the box (the .ncu-rep files stay in /tmp: gpurun brings back 64 MiB),
# the instruction-pipe breakdown of the headline kernel, and the reference arm.
mkdir -p gpurun_out
export EB200_CACHE=/tmp/eb200_cache
python bench.py --steps 3 > /dev/null 2>&1     # warm the data cache so that the profiled runs spend their time on the GPU
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/bench_ncu.log 2>&1
summ() {   # kernel regex, bench --only key, output stem, description
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$1 -s $5 -c 1 -o /tmp/$3 python bench.py --steps 1 --only $2 > gpurun_out/ncu_$3.log 2>&1
  python tools/ncu_summary.py /tmp/$3.ncu-rep gpurun_out/r02_ncu_$3_summary.txt "ncu --set full --clock-control none, $4 (python bench.py --steps 1 --only $2), round 2" > /dev/null 2>&1
}
summ k256_verify_kernel none k256_verify "k256_verify_kernel, N = 2^20" 3
ncu -i /tmp/k256_verify.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys,json
rows=list(csv.reader(sys.stdin)); h,v=rows[0],rows[2]
d=dict(zip(h,v)); rd=float(d['dram__bytes_read.sum']); wr=float(d['dram__bytes_write.sum'])
u=dict(zip(h,rows[1]))
scale={'Gbyte':1e9,'Mbyte':1e6,'Kbyte':1e3,'byte':1}
rd*=scale.get(u['dram__bytes_read.sum'],1); wr*=scale.get(u['dram__bytes_write.sum'],1)
n=1<<20
json.dump({'kernel':'k256_verify_kernel','items_in_profiled_launch':n,'dram_bytes_read':rd,'dram_bytes_write':wr,'dram_bytes_per_item':(rd+wr)/n,'dram_bytes_per_launch':rd+wr,
 'note':'ncu --set full capture of one 2^20-signature launch (python bench.py --steps 1 --only none); traffic = per-item Q table (768 B written once, partly re-read through L2) + 13 random 64-byte reads from the 436 MB fixed-base table + 161 B of inputs/outputs'}, open('gpurun_out/r02_traffic.json','w'), indent=1)"
timeout 900 ncu --clock-control none -k regex:k256_verify_kernel -s 3 -c 1 --csv --log-file gpurun_out/r02_pipes_k256_verify.csv --metrics smsp__inst_executed.sum,smsp__inst_executed_pipe_fma.sum,smsp__inst_executed_pipe_fmaheavy.sum,smsp__inst_executed_pipe_fmalite.sum,smsp__inst_executed_pipe_alu.sum,smsp__inst_executed_pipe_lsu.sum,smsp__inst_executed_pipe_cbu.sum,smsp__inst_executed_pipe_adu.sum,smsp__inst_executed_pipe_uniform.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__issue_active.avg.pct_of_peak_sustained_elapsed python bench.py --steps 1 --only none > /dev/null 2>&1
summ sw_verify_kernel p256_verify p256_verify "sw_verify_kernel<P256>, N = 2^20" 3
summ sw_verify_kernel p384_verify p384_verify "sw_verify_kernel<P384>, N = 2^20" 3
summ sw_verify_kernel p521_verify p521_verify "sw_verify_kernel<P521>, N = 2^18" 3
summ ed25519_verify_kernel ed25519_verify ed25519_verify "ed25519_verify_kernel, N = 2^20" 3
summ x25519_derive_kernel curve25519_derive x25519_derive "x25519_derive_kernel, N = 2^20" 3
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_ref_n1.json 2>&1
timeout 600 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
du -sh gpurun_out; ls gpurun_out | tr '\n' ' '
