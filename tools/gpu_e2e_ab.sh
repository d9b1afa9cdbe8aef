#!/bin/bash
# Same-box A/B of the e2e arm (host buffers through sdz_inflate_batch) for variant builds of libsdzcuda.so.
# usage: tools/gpu_e2e_ab.sh TAG variant ...   (each variant twice, interleaved)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=$1; shift
export SDZ_CORPUS_CACHE=/dev/shm/sdz_corpus
for rep in 1 2; do
  for v in "$@"; do
    SDZ_LIB=$PWD/sd-zlib_b200/csrc/variants/$v.so python bench.py --no-extras --no-checksums --cpu-sample 0 --steps 3 --warmup 3 --e2e-steps 3 2>/dev/null | tail -1 \
      | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', 'device', d['value'], 'GB/s  e2e', d['e2e']['value'], 'GB/s', d['e2e']['ms_per_step'], 'ms')" | tee -a gpurun_out/${TAG}_e2e_ab.log
  done
done
