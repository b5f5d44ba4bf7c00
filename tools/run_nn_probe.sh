timeout 300 python tools/perf_nn.py > gpurun_out/r2c_perf_nn.log 2>&1; tail -22 gpurun_out/r2c_perf_nn.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2c_nn_launches.csv python tools/perf_nn.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2c_nn_launches.csv')) if len(r)>10]
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value'); gi=h.index('Grid Size') if 'Grid Size' in h else None
for r in rows[1:60]:
    print(r[ki][:60], r[gi] if gi is not None else '', r[vi])
PY
