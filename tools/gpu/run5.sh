python -m pytest tests/test_gpu_disc.py tests/test_gpu_disc_fit.py -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['ms_per_step']); print(json.dumps(d['other_configs']['disc_reward_65536'])); print(json.dumps(d['other_configs']['disc_reward_1048576'])); print(json.dumps(d['other_configs']['a3_ppo_rollout_16384x64']))"
