#!/bin/bash
# One ncu --set full capture of the two fast-path kernels (B200_PROFILING.md recipe), after a plain run of the same command.
# 3 warm-up steps + 1 timed step launch each kernel 2 x 4 times (two chunks per step): skip the warm-up launches.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r02}
export SDZ_CORPUS_CACHE=/tmp/sdzcorpus
SHORT="python bench.py --steps 1 --warmup 3 --no-e2e --no-checksums --no-extras --cpu-sample 0"
$SHORT > gpurun_out/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"${KREGEX:-huff_tokens|lz_resolve}" -s ${KSKIP:-12} -c ${KCOUNT:-4} -o gpurun_out/${TAG}_full $SHORT > gpurun_out/${TAG}_ncu2.log 2>&1
tail -5 gpurun_out/${TAG}_ncu2.log
ls -la gpurun_out
