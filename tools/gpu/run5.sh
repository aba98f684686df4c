timeout 300 python -m pytest tests/test_gpu_disc.py -m gpu -x -q -k "one_cta" 2>&1 | tail -15
for v in 1 4; do echo "OM_DISC_VAIL2=$v"; OM_DISC_VAIL2=$v timeout 120 python tools/bench_disc.py --steps 30 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d['net'], round(d['ms']*1000,1),'us', 'frac_exec', round(d['roofline']['frac_executed'],3), 'peak', round(d['roofline']['peak'],1))
"
OM_DISC_VAIL2=$v timeout 120 python tools/bench_disc.py --steps 10 --envs 1048576 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d['net'], '1M', round(d['ms']*1000,1),'us', 'frac_exec', round(d['roofline']['frac_executed'],3))
"
done
