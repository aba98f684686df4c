for v in 1 0; do
echo "OM_DISC_VAIL2=$v"
OM_DISC_VAIL2=$v python tools/bench_disc.py --steps 30 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    print(d['net'], round(d['ms']*1000,1),'us', 'frac_exec', round(d['roofline']['frac_executed'],3), 'peak', round(d['roofline']['peak'],1))
"
OM_DISC_VAIL2=$v python tools/bench_disc.py --steps 10 --envs 1048576 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    print(d['net'], round(d['ms']*1000,1),'us', 'frac_exec', round(d['roofline']['frac_executed'],3))
"
done
