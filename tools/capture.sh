set -x
python bench.py > gpurun_out/bench_r1n.json 2> gpurun_out/bench_r1n.err || exit 1
python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/plain_r1n.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r1n.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_r1n.log 2>&1
python tools/prof_kernels.py --kernel ffn > gpurun_out/plain_ffn_n.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ffn_ln_tc2 -s 1 -c 1 -o gpurun_out/prof_ffn_r1n -f python tools/prof_kernels.py --kernel ffn > gpurun_out/ncu_ffn_n.log 2>&1
python tools/prof_kernels.py --kernel layer > gpurun_out/plain_attn_n.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:attn_cached_tc -s 12 -c 1 -o gpurun_out/prof_attn_r1n -f python tools/prof_kernels.py --kernel layer > gpurun_out/ncu_attn_n.log 2>&1
echo capture done
