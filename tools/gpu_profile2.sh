#!/bin/bash
# Profile pass (after tests / bench have run elsewhere): ncu launch list of ONE training step of the bench command
# (the first steps are skipped at full speed) and one `ncu --set full` capture of the hot kernels.
#   gpurun --timeout 1200 -- 'bash tools/gpu_profile2.sh r1d'
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
BENCH="python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-gpu-eager"
$BENCH > $OUT/${TAG}_plain_bench.log 2>&1 || { echo "plain bench failed"; tail -5 $OUT/${TAG}_plain_bench.log; exit 1; }
# model build + 3 warm-up steps come first: skip them, then list a bit more than two steps
timeout 700 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip ${SKIP:-4000} -c ${COUNT:-3600} --csv \
    --log-file $OUT/${TAG}_launches.csv $BENCH > $OUT/${TAG}_ncu_bench.log 2>&1
echo "launch list rc=$? lines=$(wc -l < $OUT/${TAG}_launches.csv)"
python tools/profile_target.py > $OUT/${TAG}_plain_target.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on \
    -k 'regex:logmel_frames|logmel_normalise|attn_fwd_tc_split_kernel|attn_bwd_tc_kernel|attn_bwd_tc_qres_kernel|gemm_gelu_kernel' -s 12 -c 8 -f -o $OUT/${TAG}_hot \
    python tools/profile_target.py > $OUT/${TAG}_ncu_hot.log 2>&1
echo "ncu full rc=$?"
ls -la $OUT | tail -8
