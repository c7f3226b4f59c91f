# kernel ms per step of a 50 Mb bench under the env given as arguments
run() { env "$@" timeout -s KILL 200 python bench.py --contig-mb 50 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$*', round(d['ms_per_step'],2), {k:round(v['ms_per_step'],2) for k,v in d['kernels'].items() if v['ms_per_step']>0}, d['clocks'])"; }
