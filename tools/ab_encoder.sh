#!/bin/bash
# A/B the observation encoder under different compile-time switches on the GPU box.
# usage: tools/ab_encoder.sh "<flags A>" "<flags B>" ...   (each arg = one PB_NVCC_EXTRA value)
# Prints step rate and encoder launch time per variant; rebuilds the default library at the end.
set -u
for flags in "$@"; do
    PB_NVCC_EXTRA="$flags" python pacbot-rl_b200/build.py --force > /dev/null || { echo "build failed: $flags"; continue; }
    for rep in 1 2; do
        timeout 300 python bench.py --steps 1000 --warmup 20 --no-cpu-baseline --no-e2e --no-dqn --no-mcts 2>/dev/null |
            python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%-60s %.1f Msteps/s  step %.4f ms  encode %.4f ms  %.0f GB/s' % ('$flags'[:60] or '(default)', d['value']/1e6, d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['achieved']))"
    done
done
PB_NVCC_EXTRA="" python pacbot-rl_b200/build.py --force > /dev/null
