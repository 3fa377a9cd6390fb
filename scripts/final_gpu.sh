#!/bin/bash
# Development aid: the round-end GPU pass in one gpurun call (tests, smoke, e2e parts, world-kernel profile, full bench line).
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/final_gputests.log 2>&1; tail -2 gpurun_out/final_gputests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
for k in 2 4; do
  python bench.py --no-train --no-cpu-baseline --steps 20 --e2e-parts $k 2>gpurun_out/final_b$k.err > gpurun_out/final_b$k.json
  python -c "import json; d=json.load(open('gpurun_out/final_b$k.json')); print('parts', d['e2e']['parts'], 'value', d['value'], 'world_ms', d['config']['single_stream']['world_kernel_ms'], 'e2e', d['e2e']['value'], 'seq', d['e2e']['sequential'], 'flags', d['config']['status_flags'])" || tail -3 gpurun_out/final_b$k.err
done
timeout 700 ncu --set full --clock-control none --import-source on -k regex:world_kernel -s 206 -c 1 -f -o gpurun_out/r02_world_v3 python bench.py --steps 10 --warmup 3 --no-train --no-cpu-baseline > gpurun_out/r02_b_under_ncu_v4.log 2>&1; tail -1 gpurun_out/r02_b_under_ncu_v4.log | cut -c1-120
