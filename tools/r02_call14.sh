#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python bench.py --rows 4000000 --legs gallery_1m --steps 10 --warmup 3 --no-cpu > gpurun_out/r02_bench_clustered.json 2> gpurun_out/r02_bench_clustered.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r02_bench_clustered.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_clustered.json') if l.startswith('{')][-1])
g=d['gallery_1m']
for k in ('1Mx768','1Mx768_clustered'):
    e=g[k]; print(k, round(e['ms_per_step'],4), round(e['value']), 'kernel ms', round(e['roofline']['avg_launch_ms'],4), e.get('parity_check',{}).get('ok'), e.get('parity_check',{}).get('product'))
PY
