# Round capture on the GPU box: bench record, launch list, ncu --set full of the two dominant kernels.
#   bash tools/capture.sh <tag> [full]      (outputs under gpurun_out/)
set -x
T=${1:-x}
python bench.py > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err || exit 1
python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/plain_$T.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_$T.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$T.log 2>&1
if [ "$2" = "full" ]; then
python tools/prof_kernels.py --kernel ffn > gpurun_out/plain_ffn_$T.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ffn_ln_tc2 -s 1 -c 1 -o gpurun_out/prof_ffn_$T -f python tools/prof_kernels.py --kernel ffn > gpurun_out/ncu_ffn_$T.log 2>&1
python tools/prof_kernels.py --kernel attn > gpurun_out/plain_attn_$T.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:attn_cached_tc -s 1 -c 1 -o gpurun_out/prof_attn_$T -f python tools/prof_kernels.py --kernel attn > gpurun_out/ncu_attn_$T.log 2>&1
fi
echo capture done
if [ "$2" = "full" ]; then
FFD_FUSED_LAYER=1 python tools/prof_layer.py 8192 > gpurun_out/plain_fused_$T.log 2>&1 && FFD_FUSED_LAYER=1 ncu --set full --clock-control none --import-source on -k regex:layer_cached -s 1 -c 1 -o gpurun_out/prof_fused_$T -f python tools/prof_layer.py 8192 > gpurun_out/ncu_fused_$T.log 2>&1
fi
