#!/bin/bash
# field-element wire format: parity, then the bench with both e2e wire formats
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "field_element or fp16_form or host_buffer" 2>&1 | tail -8
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_wire.json 2> gpurun_out/r2_bench_wire.err
tail -c 600 gpurun_out/r2_bench_wire.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_wire.json').read().strip().splitlines()[-1])
e=d['e2e']
print('value', d['value'], 'ms', d['ms_per_step'])
print('e2e fe', e['value'], e['h2d_bytes_per_step'], e['d2h_bytes_per_step'], e['frac_of_pcie'], e['pcie_roof'])
p=e['plain_arrays']
print('e2e plain', p['value'], p['frac_of_pcie'], p['pcie_roof'])
print(d.get('cpu_baseline'))
PY
