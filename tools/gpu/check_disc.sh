timeout 300 python -m pytest tests/test_gpu_disc.py tests/test_gpu_disc_fit.py -m gpu -x -q 2>&1 | tail -4
for k in 4; do for n in 65536 1048576; do
echo knob $k n $n; OM_DISC_VAIL2=$k timeout 120 python tools/bench_disc.py --envs $n --steps 30 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(json.dumps(d)[:600])"
done; done
