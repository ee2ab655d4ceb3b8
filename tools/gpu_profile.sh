#!/bin/bash
# One gpurun call: GPU parity tests, the bench line, a torch.profiler breakdown of the step, the ncu launch list of
# the bench command and one `ncu --set full` capture of the hot kernels.  Outputs land in gpurun_out/<tag>_*.
#   gpurun --timeout 1500 -- 'bash tools/gpu_profile.sh r1b'
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1

python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log
tail -3 $OUT/${TAG}_pytest.log

python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench rc=$?"; tail -c 600 $OUT/${TAG}_bench.json

python tools/bench_attn.py --bwd > $OUT/${TAG}_bench_attn.log 2>&1
cat $OUT/${TAG}_bench_attn.log

python tools/profile_step.py > $OUT/${TAG}_torch_profile.txt 2>&1
head -3 $OUT/${TAG}_torch_profile.txt

# ncu launch list of the bench command (Python-driven launches so every kernel is its own launch)
BENCH="python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-gpu-eager"
$BENCH > $OUT/${TAG}_plain_bench.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv \
    --log-file $OUT/${TAG}_launches.csv $BENCH > $OUT/${TAG}_ncu_bench.log 2>&1
echo "launch list rc=$? lines=$(wc -l < $OUT/${TAG}_launches.csv)"
if ! grep -q attn_bwd_tc_kernel $OUT/${TAG}_launches.csv; then
  # cuBLAS nvjet kernels have refused to profile before: list everything else
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv \
      -k 'regex:^(?!nvjet).*' --log-file $OUT/${TAG}_launches.csv $BENCH > $OUT/${TAG}_ncu_bench.log 2>&1
  echo "launch list (no nvjet) rc=$? lines=$(wc -l < $OUT/${TAG}_launches.csv)"
fi

# full-set capture of our hot kernels at the BASELINE shapes
python tools/profile_target.py > $OUT/${TAG}_plain_target.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on \
    -k 'regex:logmel_frames|logmel_normalise|attn_fwd_tc_kernel|attn_bwd_tc_kernel' -s 4 -c 4 -f -o $OUT/${TAG}_hot python tools/profile_target.py \
    > $OUT/${TAG}_ncu_hot.log 2>&1
echo "ncu full rc=$?"
ls -la $OUT
