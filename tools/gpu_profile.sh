#!/bin/bash
# One gpurun call: GPU parity suite, the driver's bench line, the ncu launch list and one ncu --set full capture of
# the two fast-path kernels (B200_PROFILING.md recipe).  Everything lands in gpurun_out/.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r02}
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/${TAG}_bench.json
export SDZ_CORPUS_CACHE=/tmp/sdzcorpus
SHORT="python bench.py --steps 1 --warmup 3 --no-e2e --no-checksums --no-extras --cpu-sample 0"
$SHORT > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $SHORT > gpurun_out/${TAG}_ncu1.log 2>&1
$SHORT > gpurun_out/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'huff_tokens|lz_resolve' -s 12 -c 4 -o gpurun_out/${TAG}_full $SHORT > gpurun_out/${TAG}_ncu2.log 2>&1
tail -5 gpurun_out/${TAG}_ncu2.log
ls -la gpurun_out
