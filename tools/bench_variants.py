#!/usr/bin/env python
"""Times experiment variants of libsdzcuda.so side by side on ONE GPU box (device arm of bench.py only,
full cfg3 batch, parity of the batch checked by bench.py itself).  usage: tools/bench_variants.py [--streams N] name ...
Prints one line per variant; the JSON lines land in gpurun_out/variants.jsonl."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
args = sys.argv[1:]
streams = "65536"
if args and args[0] == "--streams":
    streams, args = args[1], args[2:]
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
out = open(os.path.join(ROOT, "gpurun_out", "variants.jsonl"), "a")
for name in args:
    # name[,ENV=value...]: environment knobs of the library for this run
    name, *knobs = name.split(",")
    lib = os.path.join(ROOT, "sd-zlib_b200", "csrc", "variants", name + ".so")
    env = dict(os.environ, SDZ_LIB=lib, SDZ_CORPUS_CACHE="/dev/shm/sdz_corpus")
    for k in knobs:
        env[k.split("=")[0]] = k.split("=")[1]
    name = ",".join([name] + knobs)
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--no-checksums", "--no-e2e", "--no-extras", "--cpu-sample", "0",
                        "--streams", streams, "--steps", "5", "--warmup", "3"], env=env, capture_output=True, text=True)
    line = p.stdout.strip().splitlines()[-1] if p.stdout.strip() else ""
    try:
        j = json.loads(line)
        ph = j.get("phase_ms", {})
        print("%-16s %7.2f GB/s  kernel %.3f ms  A %.2f  B-tail %.2f  general %.2f  clocks %s" % (
            name, j["value"], j["roofline"]["kernel_ms"], ph.get("huff_tokens", 0), ph.get("lz_resolve", 0), ph.get("general_decoder", 0),
            j["clocks"]["sm_mhz"]), flush=True)
        j["variant"] = name
        out.write(json.dumps(j) + "\n")
    except Exception:
        print("%-16s FAILED rc=%d %s" % (name, p.returncode, (p.stderr or p.stdout)[-400:]), flush=True)
out.close()
