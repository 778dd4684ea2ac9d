#!/bin/bash
# end-of-round validation on a fresh box: full GPU test suite, smoke, the default bench line (config 2) with parity
o=gpurun_out
timeout 420 python -m pytest tests -m gpu -x -q > $o/r2m_pytest.log 2>&1; echo pytest_rc=$?; tail -3 $o/r2m_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 420 python bench.py --steps 5 --warmup 3 > $o/r2m_bench.json 2> $o/r2m_bench.err; echo bench_rc=$?
tail -c 300 $o/r2m_bench.err; cut -c1-260 $o/r2m_bench.json
