#!/usr/bin/env python
"""cfg4 alone: the mixed-container batch of bench.py (build_mixed_batch) on the device arm, with the library knobs given as
ENV=value arguments applied per run (each run is a fresh process of this script).  Every run checks 256 sampled streams
against the oracle.  usage: tools/bench_mixed.py [--streams N] [--dict-only] [KNOB=value[,KNOB=value] ...]
--dict-only: a batch of N zlib streams that all carry the preset dictionary (30,000 bytes of text each)."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "sd-zlib_b200", "host"))


def one(n, dict_only):
    import numpy as np, torch
    import bench as B
    from sdzlib import _native as N
    from oracle import oracle as O
    from tools import corpus as K
    ctx = N.default_context()
    threads = min(16, os.cpu_count() or 1)
    if dict_only:
        m = B.build_dict_batch(n)
    else:
        import pickle
        cache = "/dev/shm/sdz_mixed_%d.pkl" % n                         # several runs on one box share the batch
        if os.path.exists(cache):
            m = pickle.load(open(cache, "rb"))
        else:
            m = B.build_mixed_batch(n, threads, 500000)
            pickle.dump(m, open(cache, "wb"), protocol=4)
    dm = B.DeviceBatch(torch, N, m["arena"], m["off"], m["ln"], m["mode"], m["cap"], m["dict"], m["dict_len"], m["dict_adler"])
    ms, kms, ph = B.time_device_batch(ctx, dm, 3, 2)
    fs = list(ctx.last_fast_stats())
    rm, ok, _ = B.check_sampled(N, O, m, dm, n, 256)
    produced = int(sum(int(rm[i].out_len) for i in range(n)))
    print(json.dumps({"streams": n, "dict_only": dict_only, "GB/s": round(produced / (ms / 1000.0) / 1e9, 1), "ms": round(ms, 3),
                      "phase_ms[A, B after hand-over, hand-over, finalize, all]": [round(x, 3) for x in ph], "fast_path_streams": fs, "parity_ok": bool(ok)}))


if __name__ == "__main__":
    a = sys.argv[1:]
    if a and a[0] == "--one":
        one(int(a[1]), a[2] == "1")
        sys.exit(0)
    n, dict_only = 65536, False
    while a and a[0].startswith("--"):
        if a[0] == "--streams": n = int(a[1]); a = a[2:]
        elif a[0] == "--dict-only": dict_only = True; a = a[1:]
    for knobs in (a or [""]):
        env = dict(os.environ, SDZ_CORPUS_CACHE="/dev/shm/sdz_corpus")
        for k in filter(lambda x: "=" in x, knobs.split(",")):
            env[k.split("=")[0]] = k.split("=")[1]
        p = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", str(n), "1" if dict_only else "0"], env=env, capture_output=True, text=True)
        print("%-40s %s" % (knobs or "default", p.stdout.strip().splitlines()[-1] if p.stdout.strip() else "FAILED " + p.stderr[-600:]), flush=True)
