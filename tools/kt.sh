ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:lstm_tc|tail_kernel" -s 6 -c 9 --csv --log-file gpurun_out/tc6_times.csv python tests/tc_check.py full > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/tc6_times.csv')) if len(r)>10]
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value'); gi=h.index('Grid Size')
for r in rows[1:]: print(r[ki][20:52], r[gi], r[vi])
PY
timeout -s KILL 120 python tests/tc_check.py full 2>&1 | tail -2
