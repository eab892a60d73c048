#!/bin/bash
# the round-end sequence on one GPU: GPU test suite, smoke, the bench line and the reference arm
# usage (from the repo root, under gpurun): bash scripts/gpu_round.sh TAG [steps] [warmup]
tag=${1:-run}; steps=${2:-20}; warmup=${3:-5}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/${tag}_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_tests.log
tail -14 gpurun_out/${tag}_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v "done\|threads\|Render\|^$" | tail -3
timeout 1500 python bench.py --steps $steps --warmup $warmup > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
timeout 1500 python bench.py --impl reference --steps $steps --warmup $warmup > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; echo "reference rc=$?"
python - <<P
import json
d=json.load(open('gpurun_out/${tag}_bench.json')); r=json.load(open('gpurun_out/${tag}_bench_ref.json'))
print({k:d[k] for k in ('value','value_bit_identical','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], d['e2e_bit_identical'])
q=d['roofline']; print('roofline', q['kernel'], round(q['achieved']), round(q['frac'],3), 'B/item', q['algorithmic_bytes_per_item'], 'items/launch', round(q['items_per_launch']), 'traffic', q['traffic'], {a:round(b,2) for a,b in q['stage_share_of_step'].items()}, '| trace ref-layout GB/s', round(q['trace']['reference_layout_gbs']))
print('cpu_baseline', d.get('cpu_baseline',{}).get('value'), d.get('cpu_baseline',{}).get('cores'), '| reference arm', r['value'], r['config']['sample_spp_per_step'], 'spp/step', r['ms_per_step'])
print('ratio e2e/reference', d['e2e']['value']/r['value'], ' bit-identical', d['e2e_bit_identical']/r['value'])
for k,v in d['scenes'].items():
    print(k, v['workload'])
    for m in ('fast','bit_identical','s_normal'):
        if m in v: print('   ',m, {a:(round(b,2) if isinstance(b,float) else b) for a,b in v[m].items()})
    if 'roofline' in v: q=v['roofline']; print('    roofline', q['kernel'], round(q['achieved']), round(q['frac'],3), 'B/item', round(q['algorithmic_bytes_per_item']), 'launch_ms', round(q['launch_ms'],3), {a:round(b,2) for a,b in q['stage_share_of_step'].items()})
    if 'cpu_baseline' in v: print('    cpu', v['cpu_baseline']['value'], v['cpu_baseline']['cores'])
P
