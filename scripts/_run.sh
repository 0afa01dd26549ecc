timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 280 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_att.json 2> gpurun_out/bench_att.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_att.json').read())
print(d['ms_per_step'], d['value'], d['e2e']['value'])
PY
