#!/bin/bash
# stage times of blend kernels for each compile-time variant (experiment helper)
for v in 0 1 2 3; do
  SB_BWD_VARIANT=$v python bench.py --steps 15 --warmup 4 --no-cpu-baseline --no-mapping 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bwd variant $v', d['ms_per_step'], d['roofline']['stage_ms']['blend_backward'], d['roofline']['stage_ms']['blend_forward'], 'e2e', d['e2e']['ms_per_step'])"
done
SB_FWD_VARIANT=1 python bench.py --steps 15 --warmup 4 --no-cpu-baseline --no-mapping 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('fwd variant 1', d['ms_per_step'], d['roofline']['stage_ms']['blend_backward'], d['roofline']['stage_ms']['blend_forward'])"
