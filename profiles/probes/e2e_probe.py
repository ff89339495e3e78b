"""e2e breakdown on one box: the same bench.py e2e leg with the sliced-ELL kernel and with the merge-path kernel, plus the
allocator's device-allocation count per leg (a cudaMalloc inside the loop serialises the copy stream)."""
import json, os, subprocess, sys
for algo in ('auto', 'mpg'):
    env = dict(os.environ, GG_SPMM_ALGO=algo, GG_E2E_STATS='1')
    r = subprocess.run([sys.executable, 'bench.py', '--steps', '5', '--warmup', '3', '--no-cpu', '--no-comparators'], env=env,
                       capture_output=True, text=True)
    try:
        d = json.loads(r.stdout.strip().splitlines()[-1])
        print(algo, 'value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'e2e ms', d['e2e']['ms_per_step'],
              d['e2e'].get('stats'), flush=True)
    except Exception as exc:
        print(algo, 'failed', exc, r.stderr[-2000:])
