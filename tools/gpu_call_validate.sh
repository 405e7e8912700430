set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py > gpurun_out/bench_r1_validate_n1.json 2> gpurun_out/bench_r1_validate_n1.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_r1_validate_n1.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['seconds'], 'roofline', d['roofline']['frac'], 'ccsd', d['ccsd']['s_per_iter'], 'cpu', d['cpu_baseline']['value'], 'launches', d['gpu_launches'], d['clocks'])
print(open('gpurun_out/bench_r1_validate_n1.json').read().count(chr(10)), 'stdout lines')
"
