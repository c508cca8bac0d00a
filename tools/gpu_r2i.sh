#!/bin/bash
mkdir -p gpurun_out
CWG_DIRECT_EPI=1 timeout 300 python -m pytest tests/test_gpu_stages.py -x -q > gpurun_out/r2i_stages_direct.log 2>&1; rc=$?
tail -4 gpurun_out/r2i_stages_direct.log
if [ $rc -ne 0 ]; then echo "direct-epilogue stage tests failed/hung rc=$rc"; fi
CWG_DIRECT_EPI=1 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
for cfg in "CWG_EARLY_END=1 CWG_DIRECT_EPI=0" "CWG_EARLY_END=0 CWG_DIRECT_EPI=0" "CWG_EARLY_END=1 CWG_DIRECT_EPI=1" "CWG_EARLY_END=0 CWG_DIRECT_EPI=1" "CWG_EARLY_END=1 CWG_DIRECT_EPI=0"; do
  echo "== $cfg"
  env $cfg timeout 600 python bench.py --no-cpu-baseline --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('ms', round(d['ms_per_step'],2), 'frac', round(r['frac'],4), 'avg_launch', round(r['avg_launch_ms'],4), r['launch_ms_by_layer'], d['clocks']['sm_mhz'])"
done
for p in bf16x3 bf16; do
 for cfg in "CWG_DIRECT_EPI=0" "CWG_DIRECT_EPI=1"; do
  echo "== $p $cfg"
  env $cfg timeout 600 python bench.py --precision $p --no-cpu-baseline --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('ms', round(d['ms_per_step'],2), 'frac', round(r['frac'],4), 'avg_launch', round(r['avg_launch_ms'],4), r['launch_ms_by_layer'], d['clocks']['sm_mhz'])"
 done
done
