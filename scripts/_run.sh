mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest.log
timeout 280 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --dump-profile gpurun_out/perop_grp.json > gpurun_out/bench_grp.json 2> gpurun_out/bench_grp.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_grp.json').read())
print(d['ms_per_step'], d['value'], d['e2e'], d['clocks'])
PY
DISCO_TC_GRP=0 DISCO_TC_PAIR=0 timeout 280 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_old.json 2> gpurun_out/bench_old.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_old.json').read())
print(d['ms_per_step'], d['value'], d['e2e'], d['clocks'])
PY
