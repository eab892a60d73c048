#!/bin/bash
# N GPUs of one box: the multi-GPU tests (vimg_render_multi, binding) and bench.py under torchrun
# usage (under gpurun --gpus N): bash scripts/gpu_multi.sh N TAG [steps] [warmup]
n=${1:-2}; tag=${2:-multi}; steps=${3:-5}; warmup=${4:-3}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_binding.py -q > gpurun_out/${tag}_tests.log 2>&1; tail -3 gpurun_out/${tag}_tests.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps $steps --warmup $warmup > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
python - <<P
import json
d=json.load(open('gpurun_out/${tag}_bench.json'))
print({k:d[k] for k in ('value','value_bit_identical','ms_per_step','n_gpus')}, 'e2e', d['e2e']['value'], d['e2e_bit_identical'])
for k,v in d['scenes'].items():
    print(k, v['workload'], {a:(round(b,2) if isinstance(b,float) else b) for a,b in v['fast'].items()})
P
