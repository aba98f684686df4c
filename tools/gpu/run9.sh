ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.max,smsp__inst_executed.sum --cache-control none --clock-control none -k regex:"a3_|affine_scan" -s 12 -c 12 --csv --log-file gpurun_out/a3_inflow.csv python tools/bench_a3.py --steps 6 --warmup 3 > /dev/null 2>&1
grep -v "^==" gpurun_out/a3_inflow.csv | python -c "
import csv,sys
r=list(csv.DictReader(sys.stdin))
for x in r: print(x['Kernel Name'][:40], x['Metric Name'], x['Metric Value'])"
