#!/bin/bash
# A/B value vs e2e under compile-time switches: tools/ab_e2e.sh "<flags A>" "<flags B>" ...
for flags in "$@"; do
    PB_NVCC_EXTRA="$flags" python pacbot-rl_b200/build.py --force > /dev/null || { echo "build failed: $flags"; continue; }
    timeout 300 python bench.py --steps 1000 --warmup 20 --no-cpu-baseline --no-dqn --no-mcts 2>/dev/null |
        python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%-40s value %.4f ms/step  e2e %.4f ms/step' % ('$flags'[:40] or '(default)', d['ms_per_step'], 65536e3/d['e2e']['value']))"
done
PB_NVCC_EXTRA="" python pacbot-rl_b200/build.py --force > /dev/null
