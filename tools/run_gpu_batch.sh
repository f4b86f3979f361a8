mkdir -p gpurun_out
echo "=== warm launch list (cache-control none)"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 1150 -c 300 --csv --log-file gpurun_out/launches_warm.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/bench_ncu_warm.log 2>&1
tail -n 2 gpurun_out/bench_ncu_warm.log | cut -c1-200
python - <<'PY'
import csv
lines=[l for l in open('gpurun_out/launches_warm.csv') if not l.startswith('==')]
rows=list(csv.DictReader(lines))
tot=sum(float(r['Metric Value'].replace(',','')) for r in rows)
print('launches',len(rows),'sum_us',tot/1000)
PY
