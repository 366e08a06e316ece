#!/bin/bash
# what the driver runs at round end: pytest -m gpu, smoke(), bench.py (both arms)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu --timeout 900 --timeout-method=thread > gpurun_out/s_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/s_pytest.log; tail -6 gpurun_out/s_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/s_smoke.log; tail -4 gpurun_out/s_smoke.log
( time timeout 900 python bench.py ) > gpurun_out/s_bench.json 2> gpurun_out/s_bench.err; tail -c 600 gpurun_out/s_bench.json; tail -4 gpurun_out/s_bench.err
( time timeout 900 python bench.py --impl reference ) > gpurun_out/s_bench_ref.json 2> gpurun_out/s_bench_ref.err; tail -c 900 gpurun_out/s_bench_ref.json; tail -4 gpurun_out/s_bench_ref.err
