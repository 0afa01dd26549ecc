mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_conv_tc.py -x -q 2>&1 | tail -3
P="python -m disentangledcolorization_b200.tools.conv_probe"
{
timeout 60 $P --cin 64 --cout 64 --hw 256 --batch 64
timeout 60 $P --cin 64 --cout 64 --hw 256 --batch 64 --res 1
timeout 60 $P --cin 128 --cout 64 --hw 128 --batch 64 --up2 1
timeout 60 $P --cin 16 --cout 16 --hw 256 --batch 64
timeout 60 $P --cin 32 --cout 32 --hw 128 --batch 64
timeout 60 $P --cin 512 --cout 512 --hw 32 --batch 64
} 2>&1 | grep -v "^$" | tee gpurun_out/probe_alt.log
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 280 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --dump-profile gpurun_out/perop_alt.json > gpurun_out/bench_alt.json 2> gpurun_out/bench_alt.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_alt.json').read())
print(d['ms_per_step'], d['value'], d['e2e']['value'])
PY
