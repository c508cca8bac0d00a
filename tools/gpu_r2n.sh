#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_stages.py -x -q > gpurun_out/r2n_stages.log 2>&1; rc=$?
tail -3 gpurun_out/r2n_stages.log
if [ $rc -ne 0 ]; then echo "stage tests failed/hung rc=$rc"; exit 1; fi
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_bench_contract.py -x -q -m gpu 2>&1 | tail -3
for cfg in "CWG_PDL=1" "CWG_PDL=0" "CWG_PDL=1" "CWG_PDL=0"; do
  echo "== $cfg"
  env $cfg timeout 600 python bench.py --no-cpu-baseline --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('ms', round(d['ms_per_step'],2), 'e2e_ms', round(d['e2e']['ms_per_step'],2), 'frac', round(r['frac'],4), 'avg_launch', round(r['avg_launch_ms'],4), 'share', round(r['layer_share_of_step'],4), d['clocks']['sm_mhz'])"
done
