mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/pytest_gpu5.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1e.json 2> gpurun_out/bench_r1e.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
tail -4 gpurun_out/pytest_gpu5.log; tail -2 gpurun_out/smoke.log; python - <<'PY'
import json
b=json.load(open('gpurun_out/bench_r1e.json'))
print(b['value'], b['e2e']['value'], b['clocks'], b['roofline']['kernel'], b['roofline']['frac'])
PY
