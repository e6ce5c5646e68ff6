#!/bin/bash
# round 2, GPU call 2: the refactored library (graph patching, slots, new entry points) under the full GPU suite
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_call2_pytest.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/r02_call2_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_call2_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r02_call2_smoke.log
timeout 600 python bench.py --legs parity,gallery_1m --steps 20 --warmup 5 > gpurun_out/r02_bench_call2.json 2> gpurun_out/r02_bench_call2.err; echo "bench rc=$?"
tail -c 400 gpurun_out/r02_bench_call2.err
timeout 120 python tools/overhead.py > gpurun_out/r02_overhead.log 2>&1; cat gpurun_out/r02_overhead.log
