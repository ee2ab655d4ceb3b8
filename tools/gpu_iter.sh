#!/bin/bash
# quick GPU iteration: attention + layernorm parity, variant timing, timelines
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-it}
timeout 300 python -m pytest tests/test_gpu_attention_tc.py tests/test_gpu_layernorm.py -x -q 2>&1 | tail -15 | tee $OUT/${TAG}_pytest_tc.log
timeout 300 python tools/bench_variants.py 2>&1 | tee $OUT/${TAG}_variants.log
timeout 120 python tools/timeline.py bwd > $OUT/${TAG}_timeline_bwd.log 2>&1; tail -3 $OUT/${TAG}_timeline_bwd.log
timeout 120 python tools/timeline.py bwd 16 12 > $OUT/${TAG}_timeline_bwd_full.log 2>&1; tail -3 $OUT/${TAG}_timeline_bwd_full.log
timeout 120 python tools/timeline.py fwd > $OUT/${TAG}_timeline_fwd.log 2>&1; tail -3 $OUT/${TAG}_timeline_fwd.log
timeout 120 python tools/timeline.py fwd 16 12 > $OUT/${TAG}_timeline_fwd_full.log 2>&1; tail -3 $OUT/${TAG}_timeline_fwd_full.log
timeout 120 python tools/cta_log.py 2>&1 | tee $OUT/${TAG}_cta_log.log
