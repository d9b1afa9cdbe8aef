#!/bin/bash
# copy-order comparison at N GPUs: one batch on the multi-device context and the torchrun bench, with and without SDZ_H2D_FIRST
cd "$(dirname "$0")/.."
TAG=${1:-r02}; NG=${2:-8}
mkdir -p gpurun_out
export SDZ_CORPUS_CACHE=/tmp/sdzcorpus
python tools/bench_multi.py --steps 2 > gpurun_out/${TAG}_multi_default.json 2>/dev/null; cat gpurun_out/${TAG}_multi_default.json | cut -c1-260
SDZ_H2D_FIRST=1 python tools/bench_multi.py --steps 2 > gpurun_out/${TAG}_multi_h2dfirst.json 2>/dev/null; cat gpurun_out/${TAG}_multi_h2dfirst.json | cut -c1-260
SDZ_H2D_FIRST=1 SDZ_TRACE_PIPE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus $NG --steps 3 --warmup 3 --no-extras --no-checksums > gpurun_out/${TAG}_bench_n${NG}_h2dfirst.json 2> gpurun_out/${TAG}_bench_n${NG}_h2dfirst.err
python -c "import json,sys; j=json.loads(open('gpurun_out/${TAG}_bench_n${NG}_h2dfirst.json').read().strip().splitlines()[-1]); print('torchrun h2d_first e2e', j['e2e'])"
