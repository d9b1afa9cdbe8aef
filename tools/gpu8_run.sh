#!/bin/bash
# One multi-GPU gpurun call (gpurun --gpus N): host-feed ceiling of the box with N concurrent PCIe probes, the driver's bench
# under torchrun with the pipeline trace, ONE batch on a multi-device context in a single process, and the multi-device test.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r02}; NG=${2:-8}
nvidia-smi topo -m > gpurun_out/${TAG}_topo.txt 2>&1
lscpu | grep -E "NUMA|Socket|^CPU\(s\)" >> gpurun_out/${TAG}_topo.txt
for g in $(seq 0 $((NG-1))); do cat /sys/bus/pci/devices/$(nvidia-smi -i $g --query-gpu=pci.bus_id --format=csv,noheader | tr 'A-Z' 'a-z' | sed 's/^0000//')/numa_node 2>/dev/null; done | tr '\n' ' ' >> gpurun_out/${TAG}_topo.txt
# (1) ceiling: N probes at once, each 4 GiB pinned H2D / D2H / duplex
for g in $(seq 0 $((NG-1))); do CUDA_VISIBLE_DEVICES=$g python tools/pcie_probe.py > gpurun_out/${TAG}_pcie_$g.log 2>&1 & done; wait
grep -H "GB/s" gpurun_out/${TAG}_pcie_*.log > gpurun_out/${TAG}_pcie_concurrent.log; rm -f gpurun_out/${TAG}_pcie_?.log
CUDA_VISIBLE_DEVICES=0 python tools/pcie_probe.py > gpurun_out/${TAG}_pcie_alone.log 2>&1
# (2) ONE batch, one process, multi-device context
timeout 600 python tools/bench_multi.py > gpurun_out/${TAG}_multi_ctx.json 2> gpurun_out/${TAG}_multi_ctx.err; echo "multi rc=$?"; cat gpurun_out/${TAG}_multi_ctx.json
# (3) the driver's bench line at N GPUs with the pipeline trace of every rank
SDZ_TRACE_PIPE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $NG --steps 3 --warmup 3 --no-extras --no-checksums > gpurun_out/${TAG}_bench_n$NG.json 2> gpurun_out/${TAG}_bench_n$NG.err; echo "bench rc=$?"
tail -c 2500 gpurun_out/${TAG}_bench_n$NG.json
# (3b) cfg5 at N GPUs
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29534 tools/bench_large.py --gpus $NG --mib 1024 --steps 3 --warmup 1 > gpurun_out/${TAG}_large_n$NG.json 2> gpurun_out/${TAG}_large_n$NG.err; echo "large rc=$?"; tail -c 1200 gpurun_out/${TAG}_large_n$NG.json
# (4) the multi-device parity test on all devices
python -m pytest tests -m gpu -x -q -k "multi_device" 2>&1 | tail -3
