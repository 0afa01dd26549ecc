mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for pdl in 0 1 0 1; do
DISCO_TC_PDL=$pdl timeout 280 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pdl$pdl.json 2> gpurun_out/bench_pdl$pdl.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_pdl$pdl.json').read())
print($pdl, d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks'])
PY
done
