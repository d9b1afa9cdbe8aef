#!/usr/bin/env python
"""bench.py - headline benchmark: batched inflate of independent 64 KiB zlib streams.

Workload (BASELINE.json configs[2]): 65,536 synthetic-text streams of 65,536 bytes each,
deflate level 6, zlib wrapper 78 01, per GPU (weak scaling: every rank decodes its own
65,536 streams; no collective on the data path, NCCL only gathers the result records).

One step = one pass of the hot path over the whole batch:
  value    decompressed GB/s with the compressed batch already resident in HBM (device arm)
  e2e      the same through the public C ABI call sdz_inflate_batch() with pinned HOST
           buffers: staging + H2D + kernels + D2H of bytes and records inside the timed region
  roofline (compressed in + decompressed out) / decode-kernel time vs the measured HBM peak
  cpu_baseline / --impl reference: the CPU oracle (a C port of the reference's algorithm; the
           TypeScript reference itself cannot run here: no JS runtime) on all host cores.

Extra keys of the same JSON line (each parity-checked against the oracle / zlib, none of them part of `value`):
  matrix        device arm on text / binary corpora at levels 1 / 6 / 9 (north_star's measurement matrix)
  mixed_batch   BASELINE configs[3]: gzip + raw + zlib (+ preset dictionary), levels 1/6/9, stored / fixed /
                dynamic blocks, 65,536 streams per GPU
  dict_batch    65,536 zlib streams that all use a preset dictionary (InflaterOptions.dictionary)
  large_stream  BASELINE configs[4]: ONE gzip stream decoded block-parallel on all ranks (strong scaling)
  e2e_pageable  the e2e arm with separately allocated pageable inputs and a pageable output arena
                (what an N-API caller hands over)
  checksums     BASELINE configs[1]: crc32 + adler32 over 8 GiB with seed chaining, every chained value
                compared with zlib / the oracle over the full buffer
  cpu_baseline_zlib  system zlib inflate with the same threading (BASELINE.md's second CPU substitute)
"""
import argparse
import ctypes as C
import json
import os
import shutil
import subprocess
import sys
import threading
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "sd-zlib_b200", "host"))

STREAM_BYTES = 65536
N_STREAMS = 65536
LEVEL = 6


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """samples nvidia-smi during the timed region (B200_PROFILING.md clocks line)"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        mx = max([int(s[1]) for s in self.samples if s[1].isdigit()] or [0])
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples for n, v in zip(names, s[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": reasons, "samples": len(sm)}


def measured_traffic(n_streams):
    """DRAM bytes per step of the two decode kernels from this round's committed ncu capture
    (profiles/r02_traffic.json names the capture and the commit it was taken at; full-size run only)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
            t = json.load(f)
        if n_streams != N_STREAMS:
            return None, None
        return int(sum(k["dram_bytes_read"] + k["dram_bytes_write"] for k in t["kernels"].values())), t.get("capture")
    except Exception:
        return None, None


def make_corpus(n_distinct, first_index, threads, kind=None, level=LEVEL):
    from tools import corpus as K
    kind = K.TEXT if kind is None else kind
    t0 = time.time()
    cache = os.environ.get("SDZ_CORPUS_CACHE")      # tools/bench_variants.py: several runs on one box share the corpus
    path = "%s.%d.%d.%d.%d.npz" % (cache, n_distinct, first_index, kind, level) if cache else None
    if path and os.path.exists(path):
        z = np.load(path)
        return z["comp"], int(z["stride"]), z["clen"], time.time() - t0
    comp, stride, clen, _ = K.make_batch(kind, n_distinct, STREAM_BYTES, level, K.ZLIB, first_index=first_index, threads=threads)
    if path:
        np.savez(path, comp=comp, stride=np.int64(stride), clen=clen)
    return comp, stride, clen, time.time() - t0


def pack(comp, stride, clen, reps):
    """16-byte aligned packed arena of reps x distinct streams; returns (arena, off u64, len u32)."""
    nd = len(clen)
    al = (clen + np.uint64(15)) & ~np.uint64(15)
    off1 = np.zeros(nd, dtype=np.uint64)
    off1[1:] = np.cumsum(al[:-1])
    one = int(al.sum())
    arena = np.zeros(one * reps + 1024, dtype=np.uint8)
    for i in range(nd):
        o, n = int(off1[i]), int(clen[i])
        arena[o:o + n] = comp[i * stride:i * stride + n]
    for r in range(1, reps):
        arena[r * one:(r + 1) * one] = arena[:one]
    off = np.concatenate([off1 + np.uint64(r * one) for r in range(reps)])
    ln = np.tile(clen.astype(np.uint32), reps)
    return arena, off, ln


def run_reference(args, rank, world):
    """--impl reference: the CPU oracle port on all host cores, bounded sample per step."""
    if rank != 0:
        return
    from oracle import oracle as O
    cores = os.cpu_count() or 1
    n = args.ref_streams
    comp, stride, clen, _ = make_corpus(n, 0, cores)
    arena, off, ln = pack(comp, stride, clen, 1)
    out_off = np.arange(n, dtype=np.uint64) * np.uint64(STREAM_BYTES)
    out_cap = np.full(n, STREAM_BYTES, dtype=np.uint64)
    out = np.empty(n * STREAM_BYTES, dtype=np.uint8)
    modes = np.zeros(n, dtype=np.uint8)
    times = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        _, res = O.inflate_batch_mt(arena, off, ln.astype(np.uint64), out_off, out_cap, cores, modes, out)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    assert all(res[i].success for i in range(0, n, 61))
    ms = 1000.0 * sum(times) / len(times)
    gbs = n * STREAM_BYTES / (ms / 1000.0) / 1e9
    sample = "%d of the 65,536 text-64K level-6 zlib streams per step, oracle C port, %d threads" % (n, cores)
    print(json.dumps({
        "impl": "reference", "metric": "batched inflate decompressed GB/s", "value": round(gbs, 4), "unit": "GB/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "inflateBatch 65,536 x 64 KiB synthetic-text zlib streams, level 6 (bounded CPU sample)",
                   "streams_per_step": n, "stream_bytes": STREAM_BYTES},
        "cpu_baseline": {"value": round(gbs, 4), "unit": "GB/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": round(gbs, 4), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "node": shutil.which("node"),
    }))


def _i32(v):
    v &= 0xFFFFFFFF
    return v - (1 << 32) if v & 0x80000000 else v


def bench_checksums(ctx, torch, peak):
    """BASELINE configs[1]: crc32 + adler32 over an 8 GiB buffer with seed chaining (64 MiB chunks), plus a chain with
    irregular segment lengths that contain multiples of 5552.  Returns (result, verify) where verify()
    compares EVERY chained value of both chains over the full 8 GiB on the host (zlib for the bulk; the oracle - the
    reference's arithmetic, Q1 included - for every segment up to 100 MB, which covers each multiple of 5552 and the
    segment after it)."""
    import sdzlib
    from oracle import oracle as O
    total = 8 << 30
    g = torch.Generator(device="cuda")
    g.manual_seed(0x5D211B00)
    buf = torch.empty(total, dtype=torch.uint8, device="cuda")
    view = buf.view(torch.int64)
    step = 1 << 27
    for i in range(0, view.numel(), step):
        view[i:i + step].random_(generator=g)
    torch.cuda.synchronize()
    out = {}
    chunk = 64 << 20
    lens = [chunk] * (total // chunk)
    irregular = [5552 * 12000, chunk + 1, 5552, (64 << 20) - 7, 11104 * 999, 3, 5552 * 4096, 1000]
    rest = total - sum(irregular)
    irregular += [rest // 2, rest - rest // 2]
    got = {}
    for name, fn in (("crc32", sdzlib.crc32_chain), ("adler32", sdzlib.adler32_chain)):
        best = None
        for it in range(4):
            vals = fn(None, lens, ctx=ctx, device_ptr=buf.data_ptr())
            ms = ctx.last_timing()[2]
            best = ms if best is None or ms < best else best
        gbs = total / (best / 1000.0) / 1e9
        vals_irr = fn(None, irregular, ctx=ctx, device_ptr=buf.data_ptr())
        got[name] = ([int(v) for v in vals], [int(v) for v in vals_irr])
        out[name] = {"GB/s": round(gbs, 1), "ms": round(best, 3), "bytes": total, "chunks": len(lens),
                     "frac_of_hbm_peak": round(gbs / peak, 4), "final": int(vals[-1]), "final_irregular": int(vals_irr[-1])}

    def verify():
        ok = {"crc32": True, "adler32": True}
        c, a = 0, 1
        host = np.empty(chunk, dtype=np.uint8)
        ht = torch.from_numpy(host)
        for i in range(total // chunk):                                  # regular chain: all 128 values of both checksums
            ht.copy_(buf[i * chunk:(i + 1) * chunk])
            c = zlib.crc32(host, c)
            a = zlib.adler32(host, a)                                    # 2^26 is not a multiple of 5552: reference == standard
            ok["crc32"] &= _i32(c) == got["crc32"][0][i]
            ok["adler32"] &= _i32(a) == got["adler32"][0][i]
        c, a, o = 0, 1, 0
        for k, n in enumerate(irregular):                                # irregular chain: all values
            done = 0
            while done < n:
                m = min(chunk, n - done)
                ht[:m].copy_(buf[o + done:o + done + m])
                piece = host[:m]
                if n <= 100 << 20:
                    assert m == n or n > chunk
                if n <= chunk:                                           # one call, the reference's arithmetic (Q1)
                    c = O.crc32(piece.tobytes(), _i32(c))
                    a = O.adler32(piece.tobytes(), _i32(a))
                else:
                    c = zlib.crc32(piece, c & 0xFFFFFFFF)
                    a = zlib.adler32(piece, a & 0xFFFFFFFF)
                done += m
            o += n
            ok["crc32"] &= _i32(c) == got["crc32"][1][k]
            ok["adler32"] &= _i32(a) == got["adler32"][1][k]
        for name in ok:
            out[name]["parity_vs_oracle"] = bool(ok[name])
            out[name]["parity_scope"] = "all %d + %d chained values over the full 8 GiB (zlib 1.3 for the bulk, the oracle for every segment <= 64 MiB incl. each multiple of 5552)" % (len(lens), len(irregular))
        return all(ok.values())

    return out, verify, buf


def build_mixed_batch(n, threads, first_index, reps=1):
    """cfg4 at bench size: 6 corpus kinds x levels 1/6/9 x gzip / gzip+FNAME / zlib / raw, plus zlib streams with a
    preset dictionary; incompressible data stays <= 49,151 B per stream.  Streams of different kinds are
    interleaved.  Returns dict(arena, off, ln, mode, cap, dict_arena, dict_off, dict_len, dict_adler, plain_bytes)."""
    from tools import corpus as K
    from oracle import oracle as O
    kinds = [(K.TEXT, 65536), (K.BINARY, 65536), (K.TINY, 150), (K.RANDOM, 40000), (K.RUNS, 65536), (K.TEXT, 20000)]
    conts = [K.GZIP, K.GZIP_NAME, K.ZLIB, K.RAW]
    n_dict = max(16, n // 64)
    per = (n - n_dict) // (len(kinds) * 3 * len(conts))
    n_dict = n - per * len(kinds) * 3 * len(conts)
    streams = []                                                        # (bytes view, plain_len, mode, has_dict)
    fi = first_index
    for kind, plen in kinds:
        for level in (1, 6, 9):
            for cont in conts:
                comp, stride, clen, _ = K.make_batch(kind, per, plen, level, cont, first_index=fi, threads=threads)
                fi += per
                for i in range(per):
                    mode = 2 if (cont == K.RAW and i % 2) else 0
                    streams.append((comp[i * stride:i * stride + int(clen[i])], plen, mode, False))
    dic = bytes(K.generate(K.TEXT, 4242, 470))
    dictid = O.adler32(dic)
    for i in range(n_dict):
        plain = K.generate(K.TEXT, fi + i, 30000)
        s = np.frombuffer(K.compress(plain, (1, 6, 9)[i % 3], K.ZLIB_DICT, dic, dictid), dtype=np.uint8)
        streams.append((s, 30000, 1, True))
    order = np.random.RandomState(12345).permutation(len(streams))
    streams = [streams[i] for i in order]
    ln = np.array([s[0].size for s in streams], dtype=np.uint32)
    al = (ln.astype(np.uint64) + np.uint64(15)) & ~np.uint64(15)
    off = np.zeros(len(streams), dtype=np.uint64)
    off[1:] = np.cumsum(al[:-1])
    arena = np.zeros(int(al.sum()) + 1024, dtype=np.uint8)
    for i, s in enumerate(streams):
        arena[int(off[i]):int(off[i]) + s[0].size] = s[0]
    cap = np.array([(s[1] + 15) & ~15 for s in streams], dtype=np.uint32)
    mode = np.array([s[2] | (0x80 if s[3] else 0) for s in streams], dtype=np.uint8)
    dlen = np.array([len(dic) if s[3] else 0 for s in streams], dtype=np.uint32)
    dadl = np.array([dictid if s[3] else 0 for s in streams], dtype=np.int32)
    if reps > 1:                                     # multi-rank runs share the host cores: fewer distinct streams, tiled
        one = arena.size - 1024
        arena = np.concatenate([arena[:one]] * reps + [np.zeros(1024, dtype=np.uint8)])
        off = np.concatenate([off + np.uint64(r * one) for r in range(reps)])
        ln, cap, mode, dlen, dadl = (np.tile(x, reps) for x in (ln, cap, mode, dlen, dadl))
    return {"arena": arena, "off": off, "ln": ln, "mode": mode, "cap": cap, "dict": np.frombuffer(dic + b"\0" * 42, dtype=np.uint8),
            "dict_len": dlen, "dict_adler": dadl, "plain_bytes": int(sum(s[1] for s in streams)) * reps, "dictionary": dic}


def build_dict_batch(n, n_distinct=2048):
    """n zlib streams (30,000 bytes of text each, levels 1/6/9) that all name the same 470-byte preset dictionary
    (`new Inflater({dictionary})`, src/sd-inflate.ts:54-80); n_distinct different streams, tiled.  Same dict layout as
    build_mixed_batch()."""
    from tools import corpus as K
    from oracle import oracle as O
    dic = bytes(K.generate(K.TEXT, 4242, 470))
    dictid = O.adler32(dic)
    ss = [np.frombuffer(K.compress(K.generate(K.TEXT, 90000 + i, 30000), (1, 6, 9)[i % 3], K.ZLIB_DICT, dic, dictid), dtype=np.uint8)
          for i in range(min(n, n_distinct))]
    ss = [ss[i % len(ss)] for i in range(n)]
    ln = np.array([s.size for s in ss], dtype=np.uint32)
    al = (ln.astype(np.uint64) + np.uint64(15)) & ~np.uint64(15)
    off = np.zeros(n, dtype=np.uint64)
    off[1:] = np.cumsum(al[:-1])
    arena = np.zeros(int(al.sum()) + 1024, dtype=np.uint8)
    for i, s in enumerate(ss):
        arena[int(off[i]):int(off[i]) + s.size] = s
    return {"arena": arena, "off": off, "ln": ln, "mode": np.full(n, 0x81, dtype=np.uint8), "cap": np.full(n, (30000 + 15) & ~15, dtype=np.uint32),
            "dict": np.frombuffer(dic + b"\0" * 42, dtype=np.uint8), "dict_len": np.full(n, len(dic), dtype=np.uint32),
            "dict_adler": np.full(n, dictid, dtype=np.int32), "plain_bytes": 30000 * n, "dictionary": dic}


def check_sampled(N, O, m, dm, n, n_samples):
    """record + bytes of n_samples evenly spaced streams of a device batch against the oracle"""
    rm = dm.records(N)
    ok, checked = True, 0
    for i in range(0, n, max(1, n // n_samples)):
        hd = bool(m["mode"][i] & 0x80)
        exp, er = O.inflate_oneshot(m["arena"][int(m["off"][i]):int(m["off"][i]) + int(m["ln"][i])].tobytes(),
                                    dictionary=m["dictionary"] if hd else None, mode=int(m["mode"][i] & 0x7f))
        o0 = int(dm.ooff[i])
        gotb = dm.d_out[o0:o0 + int(rm[i].out_len)].cpu().numpy().tobytes()
        ok = ok and er.observable() == rm[i].observable() and (er.thrown_append or gotb == exp)
        checked += 1
    return rm, bool(ok), checked


class DeviceBatch:
    """a batch resident in HBM + the sdz_batch_dev that describes it"""

    def __init__(self, torch, N, arena, off, ln, mode=None, cap=None, dic=None, dict_len=None, dict_adler=None):
        n = len(ln)
        self.n = n
        self.d_in = torch.from_numpy(arena).cuda()
        self.d_off = torch.from_numpy(off.view(np.int64)).cuda()
        self.d_len = torch.from_numpy(ln.view(np.int32)).cuda()
        self.d_mode = torch.zeros(n, dtype=torch.uint8, device="cuda") if mode is None else torch.from_numpy(mode).cuda()
        self.d_dlen = torch.zeros(n, dtype=torch.int32, device="cuda") if dict_len is None else torch.from_numpy(dict_len.view(np.int32)).cuda()
        self.d_dadl = torch.zeros(n, dtype=torch.int32, device="cuda") if dict_adler is None else torch.from_numpy(dict_adler).cuda()
        self.d_dict = None if dic is None else torch.from_numpy(dic.copy()).cuda()
        self.d_doff = None if dic is None else torch.zeros(n, dtype=torch.int64, device="cuda")     # one shared dictionary
        cap = np.full(n, STREAM_BYTES, dtype=np.uint32) if cap is None else cap
        ooff = np.zeros(n, dtype=np.uint64)
        ooff[1:] = np.cumsum(cap[:-1].astype(np.uint64))
        self.ooff, self.cap = ooff, cap
        self.out_bytes = int(cap.astype(np.uint64).sum())
        self.d_ooff = torch.from_numpy(ooff.view(np.int64)).cuda()
        self.d_ocap = torch.from_numpy(cap.view(np.int32)).cuda()
        self.d_out = torch.empty(self.out_bytes + 64, dtype=torch.uint8, device="cuda")
        self.d_res = torch.zeros(n * C.sizeof(N.Result), dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        self.b = N.BatchDev(self.d_in.data_ptr(), self.d_off.data_ptr(), self.d_len.data_ptr(), self.d_mode.data_ptr(),
                            None if self.d_dict is None else self.d_dict.data_ptr(), None if self.d_doff is None else self.d_doff.data_ptr(),
                            self.d_dlen.data_ptr(), self.d_dadl.data_ptr(), self.d_out.data_ptr(), self.d_ooff.data_ptr(),
                            self.d_ocap.data_ptr(), self.d_res.data_ptr(), n)

    def records(self, N):
        return (N.Result * self.n).from_buffer_copy(self.d_res.cpu().numpy().tobytes())


def time_device_batch(ctx, db, steps, warmup):
    """(ms per step, decode-kernel ms, phase ms[5]) by CUDA events on the launching stream"""
    for _ in range(warmup):
        ctx.check(ctx.lib.sdz_inflate_batch_device(ctx.h, C.byref(db.b), 0, 1))
    tot = inf = 0.0
    ph = [0.0] * 5
    for _ in range(steps):
        ctx.check(ctx.lib.sdz_inflate_batch_device(ctx.h, C.byref(db.b), 0, 1))
        t = ctx.last_timing()
        inf += t[0]; tot += t[2]
        ph = [a + x for a, x in zip(ph, ctx.last_phase_timing())]
    return tot / steps, inf / steps, [x / steps for x in ph]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams", type=int, default=N_STREAMS, help="streams per GPU")
    ap.add_argument("--ref-streams", type=int, default=4096)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-sample", type=int, default=2048)
    ap.add_argument("--large-mib", type=int, default=1024, help="size of the single large stream (cfg5); 0 = skip")
    ap.add_argument("--no-checksums", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip matrix / mixed_batch / large_stream / e2e_pageable")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import sdzlib
    from sdzlib import _native as N
    from oracle import oracle as O
    from tools import corpus as K

    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = N.Context(local)
    peak, peak_kind = measured_peak()
    cores = os.cpu_count() or 1
    my_cores = max(1, cores // world)
    extras = not args.no_extras

    # ---- corpora first (host only), so that the clock sampler below sees the GPU under load and nothing else.
    # Distinct streams are generated on the host cores this rank can use; with N > 1 ranks share the cores,
    # so each rank makes streams/N distinct streams and tiles them
    n = args.streams
    reps = world if n % world == 0 else 1
    n_distinct = n // reps
    comp, stride, clen, gen_s = make_corpus(n_distinct, rank * n_distinct, my_cores)
    arena, off, ln = pack(comp, stride, clen, reps)
    del comp
    comp_bytes = int(ln.astype(np.uint64).sum())
    out_bytes = n * STREAM_BYTES
    t_extra0 = time.time()
    matrix_in = []
    mixed = None
    large = None
    if extras:
        nd8 = max(1, n // 8)
        for kind, kname in ((K.TEXT, "text"), (K.BINARY, "binary")):
            for level in (1, 6, 9):
                if kind == K.TEXT and level == LEVEL:
                    continue
                c2, s2, l2, _ = make_corpus(nd8, 700000 + rank * nd8, my_cores, kind, level)
                matrix_in.append((kname, level, pack(c2, s2, l2, n // nd8)))
        mixed = build_mixed_batch(n // reps, my_cores, 800000 + rank * n, reps)
        if args.large_mib:
            from tools import bench_large as BL
            large = BL.make_stream(args.large_mib, LEVEL)
    extra_gen_s = time.time() - t_extra0

    # ---- device-resident arm
    db = DeviceBatch(torch, N, arena, off, ln)
    d_res, d_out = db.d_res, db.d_out

    def step():
        ctx.check(ctx.lib.sdz_inflate_batch_device(ctx.h, C.byref(db.b), 0, 1))
        t = ctx.last_timing() + ctx.last_phase_timing()
        if world > 1:   # K8: gather the fixed-size records (never payload) to every rank
            parts = [torch.empty_like(d_res) for _ in range(world)]
            dist.all_gather(parts, d_res)
            torch.cuda.synchronize()
        return t

    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(args.warmup):
        step()
    # parity of the timed configuration: records + sampled bytes vs the oracle
    recs = db.records(N)
    bad = sum(1 for i in range(n) if not (recs[i].success and recs[i].checksum_state == 1 and recs[i].out_len == STREAM_BYTES))
    sample_ids = list(range(0, n, max(1, n // 16)))[:16]
    for i in sample_ids:
        o, l = int(off[i]), int(ln[i])
        exp, er = O.inflate_oneshot(arena[o:o + l].tobytes())
        got = d_out[i * STREAM_BYTES:(i + 1) * STREAM_BYTES].cpu().numpy().tobytes()
        if got != exp or er.observable() != recs[i].observable():
            bad += 1
    if bad:
        raise SystemExit("parity failure in the benchmark batch: %d streams" % bad)

    launches0 = ctx.launch_count()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t_wall0 = time.perf_counter()
    k_inf = k_fin = k_tot = 0.0
    k_phase = [0.0] * 5
    for _ in range(args.steps):
        t = step()
        k_inf += t[0]; k_fin += t[1]; k_tot += t[2]
        k_phase = [a + x for a, x in zip(k_phase, t[3:8])]
    k_phase = [x / args.steps for x in k_phase]
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall_ms = (time.perf_counter() - t_wall0) * 1000.0 / args.steps
    launches = ctx.launch_count() - launches0
    fast_stats = list(ctx.last_fast_stats())

    dev_ms = k_tot / args.steps                      # CUDA events on the launching stream
    step_ms = wall_ms if world > 1 else dev_ms       # multi-rank: includes the record gather
    if world > 1:
        tt = torch.tensor([step_ms, dev_ms, k_inf / args.steps], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        step_ms, dev_ms, inf_ms = [float(x) for x in tt.tolist()]
        tot = torch.tensor([float(out_bytes), float(comp_bytes)], device="cuda", dtype=torch.float64)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        all_out, all_comp = [float(x) for x in tot.tolist()]
    else:
        inf_ms = k_inf / args.steps
        all_out, all_comp = float(out_bytes), float(comp_bytes)

    value = all_out / (step_ms / 1000.0) / 1e9
    achieved = (comp_bytes + out_bytes) / (inf_ms / 1000.0) / 1e9     # per GPU, the decode kernels

    def reduce_max(x):
        if world == 1:
            return x
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def reduce_sum(x):
        if world == 1:
            return x
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- north_star's matrix: text / binary x levels 1 / 6 / 9, device arm
    matrix = None
    if extras:
        matrix = {"text_l6": {"GB/s": round(value, 1), "frac": round(achieved / peak, 4), "kernel_ms": round(inf_ms, 3)}}
        del db
        for kname, level, (a2, o2, l2) in matrix_in:
            d2 = DeviceBatch(torch, N, a2, o2, l2)
            ms, kms, _ = time_device_batch(ctx, d2, 3, 2)
            r2 = d2.records(N)
            ok = all(r2[i].success and r2[i].checksum_state == 1 and r2[i].out_len == STREAM_BYTES for i in range(n))
            for i in (0, n // 3, n - 1):
                exp, er = O.inflate_oneshot(a2[int(o2[i]):int(o2[i]) + int(l2[i])].tobytes())
                ok = ok and d2.d_out[i * STREAM_BYTES:(i + 1) * STREAM_BYTES].cpu().numpy().tobytes() == exp and er.observable() == r2[i].observable()
            cb = int(l2.astype(np.uint64).sum())
            ms, kms = reduce_max(ms), reduce_max(kms)
            matrix["%s_l%d" % (kname, level)] = {"GB/s": round(reduce_sum(float(out_bytes)) / (ms / 1000.0) / 1e9, 1), "kernel_ms": round(kms, 3),
                                                "frac": round((cb + out_bytes) / (kms / 1000.0) / 1e9 / peak, 4),
                                                "compressed_bytes_per_gpu": cb, "parity_ok": bool(ok)}
            del d2
        matrix_in = None
        torch.cuda.empty_cache()

    # ---- cfg4: the mixed-container batch
    mixed_out = None
    if extras:
        m = mixed
        dm = DeviceBatch(torch, N, m["arena"], m["off"], m["ln"], m["mode"], m["cap"], m["dict"], m["dict_len"], m["dict_adler"])
        ms, kms, ph = time_device_batch(ctx, dm, 3, 2)
        fs = list(ctx.last_fast_stats())
        rm, ok, checked = check_sampled(N, O, m, dm, n, 256)
        produced = int(sum(int(rm[i].out_len) for i in range(n)))
        cb = int(m["ln"].astype(np.uint64).sum())
        ms, kms = reduce_max(ms), reduce_max(kms)
        mixed_out = {"workload": "cfg4: gzip / gzip+FNAME / zlib / raw / zlib+dictionary, levels 1/6/9, text / binary / tiny (fixed) / incompressible (stored) / runs",
                     "streams_per_gpu": n, "GB/s": round(reduce_sum(float(produced)) / (ms / 1000.0) / 1e9, 1), "ms": round(ms, 3), "kernel_ms": round(kms, 3),
                     "frac": round((cb + produced) / (kms / 1000.0) / 1e9 / peak, 4), "compressed_bytes_per_gpu": cb, "out_bytes_per_gpu": produced,
                     "fast_path_streams": fs, "phase_ms": [round(x, 3) for x in ph], "n_gpus": world,
                     "parity_ok": bool(ok), "parity_scope": "%d sampled streams: record + bytes vs the oracle" % checked}
        del dm
        mixed = None
        torch.cuda.empty_cache()

    # ---- streams with a preset dictionary (InflaterOptions.dictionary): finished by the two-phase path since round 2
    dict_out = None
    if extras:
        m = build_dict_batch(n)
        dm = DeviceBatch(torch, N, m["arena"], m["off"], m["ln"], m["mode"], m["cap"], m["dict"], m["dict_len"], m["dict_adler"])
        ms, kms, ph = time_device_batch(ctx, dm, 3, 2)
        fs = list(ctx.last_fast_stats())
        rm, ok, checked = check_sampled(N, O, m, dm, n, 64)
        produced = int(sum(int(rm[i].out_len) for i in range(n)))
        cb = int(m["ln"].astype(np.uint64).sum())
        ms, kms = reduce_max(ms), reduce_max(kms)
        dict_out = {"workload": "%d zlib streams x 30,000 B text, levels 1/6/9, all with FDICT + the caller's 470-byte preset dictionary" % n,
                    "GB/s": round(reduce_sum(float(produced)) / (ms / 1000.0) / 1e9, 1), "ms": round(ms, 3), "kernel_ms": round(kms, 3),
                    "frac": round((cb + produced) / (kms / 1000.0) / 1e9 / peak, 4), "fast_path_streams": fs,
                    "phase_ms": [round(x, 3) for x in ph], "parity_ok": bool(ok), "parity_scope": "%d sampled streams: record + bytes vs the oracle" % checked}
        del dm, m
        torch.cuda.empty_cache()

    # ---- cfg5: one large gzip stream on all ranks
    large_out = None
    if extras and large is not None:
        from tools import bench_large as BL

        class A0:
            steps, warmup = 3, 1
        stream, crc = large
        n_out = args.large_mib << 20
        r = BL.single_gpu(A0, stream, crc, n_out, ctx) if world == 1 else BL.multi_gpu(A0, stream, crc, n_out, ctx)
        if r is not None:
            large_out = {"workload": "cfg5: one %d MiB gzip stream (level %d, %d blocks), two-pass block-parallel decode" % (args.large_mib, LEVEL, r["blocks"]),
                         "ms": r["ms_per_step"], "GB/s": r["value"], "n_gpus": world, "scaling": "strong",
                         "frac": round((len(stream) + n_out) / (r["ms_per_step"] / 1000.0) / 1e9 / (peak * world), 4),
                         "compressed_bytes": len(stream), "e2e": r["e2e"], "gpu_launches": r["gpu_launches"],
                         "parity_ok": True, "parity_scope": "record success + CRC-32 of all decoded bytes vs zlib (asserted)",
                         "timing": "best of 3, host clock around the whole call, max over ranks"}
        large = None
        torch.cuda.empty_cache()

    # ---- e2e arm: host buffers through the public C ABI
    e2e = None
    e2e_pageable = None
    if not args.no_e2e:
        lib = ctx.lib
        h_in = lib.sdz_host_alloc_near(ctx.h, arena.size)
        h_out = lib.sdz_host_alloc_near(ctx.h, out_bytes)
        C.memmove(h_in, arena.ctypes.data, arena.size)
        ins = (N.In * n)()
        for i in range(n):
            ins[i].data = h_in + int(off[i]); ins[i].len = int(ln[i]); ins[i].mode = 0
        o_off = (np.arange(n, dtype=np.uint64) * np.uint64(STREAM_BYTES))
        o_cap = np.full(n, STREAM_BYTES, dtype=np.uint64)
        hres = (N.Result * n)()
        # Ceiling of this arm on THIS box: the call has to bring out_bytes back over the link, so it cannot beat the rate of a
        # plain pinned device -> host copy (1 GiB, best of 3; all ranks copy at the same time, as they do in the timed call)
        probe = min(1 << 30, out_bytes)
        d_probe = lib.sdz_device_alloc(ctx.h, probe)
        d2h_gbs = h2d_gbs = 0.0
        if d_probe:
            for fn, which in ((lib.sdz_memcpy_h2d, "h2d"), (lib.sdz_memcpy_d2h, "d2h")):
                best = 1e9
                for _ in range(3):
                    if world > 1:
                        dist.barrier()
                    t0 = time.perf_counter()
                    ctx.check(fn(ctx.h, d_probe, h_out, probe) if which == "h2d" else fn(ctx.h, h_out, d_probe, probe))
                    best = min(best, time.perf_counter() - t0)
                if which == "h2d":
                    h2d_gbs = probe / best / 1e9
                else:
                    d2h_gbs = probe / best / 1e9
            lib.sdz_device_free(ctx.h, d_probe)
        times = []
        for it in range(1 + args.e2e_steps):
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            ctx.check(lib.sdz_inflate_batch(ctx.h, ins, n, h_out, o_off.ctypes.data, o_cap.ctypes.data, hres, 0))
            dt = time.perf_counter() - t0
            if it:
                times.append(dt)
        e_ms = reduce_max(1000.0 * sum(times) / len(times))
        ok = all(hres[i].success for i in range(0, n, 97))
        first = (C.c_uint8 * STREAM_BYTES).from_address(h_out)
        exp0, _ = O.inflate_oneshot(arena[int(off[0]):int(off[0]) + int(ln[0])].tobytes())
        ok = ok and bytes(first) == exp0
        sum_d2h, sum_h2d = reduce_sum(d2h_gbs), reduce_sum(h2d_gbs)
        e2e = {"value": round(all_out / (e_ms / 1000.0) / 1e9, 3), "unit": "GB/s", "ms_per_step": round(e_ms, 2),
               "h2d_bytes_per_step": int(arena.size + n * 41), "d2h_bytes_per_step": int(out_bytes + n * C.sizeof(N.Result)),
               "parity_ok": bool(ok), "numa_node": int(lib.sdz_ctx_numa_node(ctx.h)),
               "link_d2h_gbs": round(sum_d2h, 1), "link_h2d_gbs": round(sum_h2d, 1),
               "ceiling_gbs": round(sum_d2h * all_out / (world * float(out_bytes + n * C.sizeof(N.Result))), 1) if sum_d2h > 0 else None,
               "ceiling": "decompressed GB/s if the device -> host bytes of the call moved at the measured pinned-copy rate of this box (sum over ranks, all copying at once) and nothing else took time",
               "api": "sdz_inflate_batch: host pointers into one pinned arena (zero-copy DMA), 4,096-stream sub-batches pipelined over copy streams and 3 compute lanes"}
        if e2e["ceiling_gbs"]:
            e2e["frac_of_ceiling"] = round(e2e["value"] / e2e["ceiling_gbs"], 3)
        lib.sdz_host_free(h_in); lib.sdz_host_free(h_out)
        if extras:
            # what an N-API caller hands over: one pageable allocation per input buffer, a pageable output arena
            bufs = [arena[int(off[i]):int(off[i]) + int(ln[i])].copy() for i in range(n)]
            for i in range(n):
                ins[i].data = bufs[i].ctypes.data
            p_out = np.empty(out_bytes, dtype=np.uint8)
            times = []
            for it in range(2):
                if world > 1:
                    dist.barrier()
                t0 = time.perf_counter()
                ctx.check(lib.sdz_inflate_batch(ctx.h, ins, n, p_out.ctypes.data, o_off.ctypes.data, o_cap.ctypes.data, hres, 0))
                dt = time.perf_counter() - t0
                if it:
                    times.append(dt)
            p_ms = reduce_max(1000.0 * sum(times) / len(times))
            ok = all(hres[i].success for i in range(0, n, 97)) and p_out[:STREAM_BYTES].tobytes() == exp0
            e2e_pageable = {"value": round(all_out / (p_ms / 1000.0) / 1e9, 3), "unit": "GB/s", "ms_per_step": round(p_ms, 2), "parity_ok": bool(ok),
                            "api": "sdz_inflate_batch: %d separately allocated pageable input buffers (staged by host threads into pinned memory), pageable output arena" % n}
            del bufs, p_out

    checks = None
    verify = None
    if rank == 0 and world == 1 and not args.no_checksums:
        d_out = d_res = None
        torch.cuda.empty_cache()
        checks, verify, ck_buf = bench_checksums(ctx, torch, peak)

    sampler.stop_flag = True
    sampler.join(timeout=3)

    if verify is not None:
        verify()                                     # host work: full-buffer comparison of every chained value
        del ck_buf
        torch.cuda.empty_cache()

    # ---- CPU baselines (rank 0, N = 1 only): the oracle port and system zlib on a bounded sample, all host cores
    cpu = cpu_zlib = None
    if rank == 0 and world == 1 and args.cpu_sample > 0:
        ns = min(args.cpu_sample, n)
        o_off = np.arange(ns, dtype=np.uint64) * np.uint64(STREAM_BYTES)
        o_cap = np.full(ns, STREAM_BYTES, dtype=np.uint64)
        best = None
        for _ in range(2):
            t0 = time.perf_counter()
            O.inflate_batch_mt(arena, off[:ns], ln[:ns].astype(np.uint64), o_off, o_cap, cores, np.zeros(ns, dtype=np.uint8))
            dt = time.perf_counter() - t0
            best = dt if best is None or dt < best else best
        cpu = {"value": round(ns * STREAM_BYTES / best / 1e9, 4), "unit": "GB/s", "cores": cores, "kind": "port",
               "sample": "%d of the %d streams, oracle C port of the reference algorithm, %d threads, best of 2" % (ns, n, cores)}
        from concurrent.futures import ThreadPoolExecutor
        views = [arena[int(off[i]):int(off[i]) + int(ln[i])] for i in range(ns)]
        best = None
        with ThreadPoolExecutor(cores) as ex:
            for _ in range(2):
                t0 = time.perf_counter()
                total = sum(ex.map(lambda v: len(zlib.decompress(v)), views, chunksize=16))
                dt = time.perf_counter() - t0
                best = dt if best is None or dt < best else best
        assert total == ns * STREAM_BYTES
        cpu_zlib = {"value": round(ns * STREAM_BYTES / best / 1e9, 4), "unit": "GB/s", "cores": cores, "kind": "system zlib %s (not the reference's algorithm)" % zlib.ZLIB_RUNTIME_VERSION,
                    "sample": "%d of the %d streams, zlib.decompress in %d threads, best of 2" % (ns, n, cores)}

    if rank == 0:
        traffic, traffic_src = measured_traffic(n)
        print(json.dumps({
            "metric": "batched inflate decompressed GB/s", "value": round(value, 2), "unit": "GB/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(step_ms, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "inflateBatch 65,536 x 64 KiB synthetic-text zlib streams, level 6, per GPU",
                       "streams_per_gpu": n, "distinct_streams_per_gpu": n_distinct, "stream_bytes": STREAM_BYTES,
                       "compressed_bytes_per_gpu": comp_bytes, "l2": "inputs+outputs (%.1f GB) far exceed the 126 MB L2" % ((comp_bytes + out_bytes) / 1e9),
                       "corpus_gen_s": round(gen_s, 1), "extra_corpora_gen_s": round(extra_gen_s, 1)},
            "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 4), "traffic": traffic, "traffic_source": traffic_src, "peak_kind": peak_kind,
                         "kernel": "huff_tokens_kernel + lz_resolve2_kernel (phase A / phase B, overlapped on two streams; + hand-over run of inflate_kernel)",
                         "kernel_ms": round(inf_ms, 3), "finalize_ms": round(k_fin / args.steps, 3),
                         "algorithmic_bytes": comp_bytes + out_bytes},
            "cpu_baseline": cpu,
            "cpu_baseline_zlib": cpu_zlib,
            "node": shutil.which("node"),
            "e2e": e2e,
            "e2e_pageable": e2e_pageable,
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
            "device_ms_per_step": round(dev_ms, 3),
            "phase_ms": {"huff_tokens": round(k_phase[0], 3), "lz_resolve": round(k_phase[1] + k_phase[2], 3), "general_decoder": round(k_phase[2], 3),
                         "finalize": round(k_phase[3], 3), "fast_path_streams": fast_stats,
                         "note": "huff_tokens = phase A of all chunks (phase B of the earlier chunks runs next to it); lz_resolve = what is left of "
                                 "phase B after the last phase A; general_decoder = the hand-over run, queued behind the last phase A and running NEXT TO "
                                 "that rest of phase B (it is part of lz_resolve's span; with an empty list it only waits for a free SM slot)"},
            "matrix": matrix,
            "mixed_batch": mixed_out,
            "dict_batch": dict_out,
            "large_stream": large_out,
            "checksums": checks,
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
