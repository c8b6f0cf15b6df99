set -x
python -m pytest tests/test_dqn_gpu.py -x -q -k "clip_adam or short_training or target_and_loss" 2>&1 | tail -3
python tools/small_encode_latency.py 2>&1 | tail -1
PB_SMALL_CARVEOUT=0 python tools/small_encode_latency.py 2>&1 | tail -1
PB_SMALL_ENC_MAX=0 python tools/small_encode_latency.py 2>&1 | tail -1
cp pacbot-rl_b200/libpacbot_b200.so /tmp/lib_default.so
PB_NVCC_EXTRA="-DPB_OBS_SMALL_THREADS=256" python pacbot-rl_b200/build.py --force > /dev/null && python tools/small_encode_latency.py 2>&1 | tail -1
PB_NVCC_EXTRA="-DPB_OBS_SMALL_THREADS=1024" python pacbot-rl_b200/build.py --force > /dev/null && python tools/small_encode_latency.py 2>&1 | tail -1
cp /tmp/lib_default.so pacbot-rl_b200/libpacbot_b200.so
python tools/dropin_latency.py 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 --no-mcts > gpurun_out/s4_bench4.json 2> gpurun_out/s4_bench4.err; echo "bench exit $?"
