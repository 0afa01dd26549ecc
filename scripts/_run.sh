mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_conv_tc.py -x -q 2>&1 | tail -3
timeout 280 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --dump-profile gpurun_out/perop_c2.json > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
tail -3 gpurun_out/bench_c2.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_c2.json').read())
print(d['ms_per_step'], d['value'], d['e2e'])
PY
DISCO_TC_CONST=0 timeout 280 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c0.json 2> gpurun_out/bench_c0.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_c0.json').read())
print(d['ms_per_step'], d['value'], d['e2e'])
PY
