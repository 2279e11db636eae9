#!/bin/bash
# whole GPU suite + smoke + default bench of both arms (what the driver runs at round end)
set -u
O=gpurun_out/r2full
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -rA > $O/pytest_full.log 2>&1
echo "pytest rc=$?" >> $O/pytest_full.log
grep -E "passed|failed" $O/pytest_full.log | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log; tail -2 $O/smoke.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_ours.json 2> $O/bench_ours.err; echo "ours rc=$?"
tail -n 3 $O/bench_ours.err
