#!/bin/bash
# A/B of the rollout pipeline switches on one box, device timeline of the 20-step protocol each:
#   PB_STEP_CARVEOUT=0|1       k_step prefers the encoder's shared-memory carve-out (default 1)
#   PB_ROLL_ENC_STREAMS=1|2    consecutive encoders on one stream / on two alternating ones (default 2)
# usage: tools/ab_rollout.sh > gpurun_out/ab_rollout.txt
for carve in 0 1; do
  for streams in 1 2; do
    echo "== PB_STEP_CARVEOUT=$carve PB_ROLL_ENC_STREAMS=$streams"
    PB_STEP_CARVEOUT=$carve PB_ROLL_ENC_STREAMS=$streams python tools/timeline.py --steps 20 --warmup 5 --repeat 4 \
      --summary-only --out gpurun_out/r02_timeline_c${carve}_s${streams}.json | \
      python -c "
import sys, json
for line in sys.stdin:
    r = json.loads(line)
    print('  ms/step %.4f  region %.1f us  head %.1f  enc mean %.1f [min %.1f max %.1f]  gap mean %.2f  tail %.1f  k_step %.1f' % (
        r['event_ms_per_step'], r['region_us'], r['head_us'], r['encoder_us_mean'], r['encoder_us_min_max'][0],
        r['encoder_us_min_max'][1], r['gap_us_mean'], r['tail_us'], r['k_step_us_mean']))
"
  done
done
