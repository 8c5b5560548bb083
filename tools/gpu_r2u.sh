#!/bin/bash
# full GPU suite, smoke, then the driver's bench command
O=gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > $O/tests_r2u.log 2>&1; tail -4 $O/tests_r2u.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_r2u.log 2>&1; tail -2 $O/smoke_r2u.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench1024_r2u.json 2> $O/bench1024_r2u.err; tail -c 300 $O/bench1024_r2u.err
