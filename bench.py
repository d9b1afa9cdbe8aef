#!/usr/bin/env python
"""bench.py - headline benchmark: batched inflate of independent 64 KiB zlib streams.

Workload (BASELINE.json configs[2]): 65,536 synthetic-text streams of 65,536 bytes each,
deflate level 6, zlib wrapper 78 01, per GPU (weak scaling: every rank decodes its own
65,536 streams; no collective on the data path, NCCL only gathers the result records).

One step = one pass of the hot path over the whole batch:
  value    decompressed GB/s with the compressed batch already resident in HBM (device arm)
  e2e      the same through the public C ABI call sdz_inflate_batch() with pinned HOST
           buffers: staging + H2D + kernels + D2H of bytes and records inside the timed region
  roofline (compressed in + decompressed out) / inflate-kernel time vs the measured HBM peak
  cpu_baseline / --impl reference: the CPU oracle (a C port of the reference's algorithm; the
           TypeScript reference itself cannot run here: no JS runtime) on all host cores.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

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


def measured_traffic(kernel, n_streams):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (full-size run only)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            t = json.load(f)[kernel]
        return int(t["dram_bytes_read"] + t["dram_bytes_write"]) if n_streams == N_STREAMS else None
    except Exception:
        return None


def make_corpus(n_distinct, first_index, threads):
    from tools import corpus as K
    t0 = time.time()
    cache = os.environ.get("SDZ_CORPUS_CACHE")      # tools/bench_variants.py: several runs on one box share the corpus
    path = "%s.%d.%d.npz" % (cache, n_distinct, first_index) if cache else None
    if path and os.path.exists(path):
        z = np.load(path)
        return z["comp"], int(z["stride"]), z["clen"], time.time() - t0
    comp, stride, clen, _ = K.make_batch(K.TEXT, n_distinct, STREAM_BYTES, LEVEL, K.ZLIB, first_index=first_index, threads=threads)
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
    }))


def bench_checksums(ctx, torch, peak):
    """BASELINE configs[1]: crc32 + adler32 over an 8 GiB buffer with seed chaining (64 MiB chunks)."""
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
    irregular = [5552 * 12000, chunk + 1, 5552, (64 << 20) - 7, 11104 * 999, 3, 5552 * 4096]
    rest = total - sum(irregular)
    irregular += [rest // 2, rest - rest // 2]
    for name, fn in (("crc32", sdzlib.crc32_chain), ("adler32", sdzlib.adler32_chain)):
        best = None
        for it in range(4):
            vals = fn(None, lens, ctx=ctx, device_ptr=buf.data_ptr())
            ms = ctx.last_timing()[2]
            best = ms if best is None or ms < best else best
        gbs = total / (best / 1000.0) / 1e9
        vals_irr = fn(None, irregular, ctx=ctx, device_ptr=buf.data_ptr())
        # parity on a bounded part: first two 64 MiB chunks and the first irregular segments vs the oracle
        host = buf[:2 * chunk].cpu().numpy()
        ofn = O.crc32 if name == "crc32" else O.adler32
        seed0 = 0 if name == "crc32" else 1
        v0 = ofn(host[:chunk].tobytes(), seed0)
        v1 = ofn(host[chunk:2 * chunk].tobytes(), v0)
        ok = int(vals[0]) == v0 and int(vals[1]) == v1
        hi = buf[:irregular[0] + irregular[1] + irregular[2]].cpu().numpy()
        s, o = seed0, 0
        for k in range(3):
            s = ofn(hi[o:o + irregular[k]].tobytes(), s)
            o += irregular[k]
            ok = ok and int(vals_irr[k]) == s
        out[name] = {"GB/s": round(gbs, 1), "ms": round(best, 3), "bytes": total, "chunks": len(lens),
                     "frac_of_hbm_peak": round(gbs / peak, 4), "parity_vs_oracle": bool(ok),
                     "final": int(vals[-1]), "final_irregular": int(vals_irr[-1])}
    del buf
    torch.cuda.empty_cache()
    return out


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
    ap.add_argument("--no-checksums", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
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

    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = N.Context(local)
    peak, peak_kind = measured_peak()
    cores = os.cpu_count() or 1

    # ---- corpus: distinct streams are generated on the host cores this rank can use; with
    # N > 1 ranks share the cores, so each rank makes streams/N distinct streams and tiles them
    n = args.streams
    reps = world if n % world == 0 else 1
    n_distinct = n // reps
    comp, stride, clen, gen_s = make_corpus(n_distinct, rank * n_distinct, max(1, cores // world))
    arena, off, ln = pack(comp, stride, clen, reps)
    comp_bytes = int(ln.astype(np.uint64).sum())
    out_bytes = n * STREAM_BYTES

    # ---- device-resident arm
    d_in = torch.from_numpy(arena).cuda()
    d_off = torch.from_numpy(off.view(np.int64)).cuda()
    d_len = torch.from_numpy(ln.view(np.int32)).cuda()
    d_mode = torch.zeros(n, dtype=torch.uint8, device="cuda")                  # SDZ_MODE_SNIFF, as inflate() does
    d_dlen = torch.zeros(n, dtype=torch.int32, device="cuda")
    d_dadl = torch.zeros(n, dtype=torch.int32, device="cuda")
    d_ooff = (torch.arange(n, dtype=torch.int64, device="cuda") * STREAM_BYTES)
    d_ocap = torch.full((n,), STREAM_BYTES, dtype=torch.int32, device="cuda")
    d_out = torch.empty(out_bytes + 64, dtype=torch.uint8, device="cuda")
    d_res = torch.zeros(n * C.sizeof(N.Result), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    b = N.BatchDev(d_in.data_ptr(), d_off.data_ptr(), d_len.data_ptr(), d_mode.data_ptr(), None, None,
                   d_dlen.data_ptr(), d_dadl.data_ptr(), d_out.data_ptr(), d_ooff.data_ptr(), d_ocap.data_ptr(),
                   d_res.data_ptr(), n)

    def step():
        ctx.check(ctx.lib.sdz_inflate_batch_device(ctx.h, C.byref(b), 0, 1))
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
    # parity of the timed configuration: records + a checksum of checksums + sampled bytes vs the oracle
    res_host = d_res.cpu().numpy()
    recs = (N.Result * n).from_buffer_copy(res_host.tobytes())
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
    achieved = (comp_bytes + out_bytes) / (inf_ms / 1000.0) / 1e9     # per GPU, dominant kernel

    # ---- e2e arm: host buffers through the public C ABI
    e2e = None
    if not args.no_e2e:
        lib = ctx.lib
        h_in = lib.sdz_host_alloc(arena.size)
        h_out = lib.sdz_host_alloc(out_bytes)
        C.memmove(h_in, arena.ctypes.data, arena.size)
        ins = (N.In * n)()
        for i in range(n):
            ins[i].data = h_in + int(off[i]); ins[i].len = int(ln[i]); ins[i].mode = 0
        o_off = (np.arange(n, dtype=np.uint64) * np.uint64(STREAM_BYTES))
        o_cap = np.full(n, STREAM_BYTES, dtype=np.uint64)
        hres = (N.Result * n)()
        times = []
        for it in range(1 + args.e2e_steps):
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            ctx.check(lib.sdz_inflate_batch(ctx.h, ins, n, h_out, o_off.ctypes.data, o_cap.ctypes.data, hres, 0))
            dt = time.perf_counter() - t0
            if it:
                times.append(dt)
        e_ms = 1000.0 * sum(times) / len(times)
        ok = all(hres[i].success for i in range(0, n, 97))
        first = (C.c_uint8 * STREAM_BYTES).from_address(h_out)
        exp0, _ = O.inflate_oneshot(arena[int(off[0]):int(off[0]) + int(ln[0])].tobytes())
        ok = ok and bytes(first) == exp0
        if world > 1:
            tt = torch.tensor([e_ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            e_ms = float(tt.item())
        e2e = {"value": round(all_out / (e_ms / 1000.0) / 1e9, 3), "unit": "GB/s", "ms_per_step": round(e_ms, 2),
               "h2d_bytes_per_step": int(arena.size + n * 41), "d2h_bytes_per_step": int(out_bytes + n * C.sizeof(N.Result)),
               "parity_ok": bool(ok), "api": "sdz_inflate_batch: host pointers into one pinned arena (zero-copy DMA), 4,096-stream sub-batches pipelined over copy streams and 3 compute lanes"}
        lib.sdz_host_free(h_in); lib.sdz_host_free(h_out)

    sampler.stop_flag = True
    sampler.join(timeout=3)

    # ---- CPU baseline (rank 0, N = 1 only): the oracle port on a bounded sample
    cpu = None
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

    checks = None
    if rank == 0 and world == 1 and not args.no_checksums:
        del d_out
        torch.cuda.empty_cache()
        checks = bench_checksums(ctx, torch, peak)

    if rank == 0:
        print(json.dumps({
            "metric": "batched inflate decompressed GB/s", "value": round(value, 2), "unit": "GB/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(step_ms, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "inflateBatch 65,536 x 64 KiB synthetic-text zlib streams, level 6, per GPU",
                       "streams_per_gpu": n, "distinct_streams_per_gpu": n_distinct, "stream_bytes": STREAM_BYTES,
                       "compressed_bytes_per_gpu": comp_bytes, "l2": "inputs+outputs (%.1f GB) far exceed the 126 MB L2" % ((comp_bytes + out_bytes) / 1e9),
                       "lanes_per_stream": int(os.environ.get("SDZ_GROUP", "4")), "corpus_gen_s": round(gen_s, 1)},
            "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 4), "traffic": measured_traffic("inflate_kernel", n), "peak_kind": peak_kind,
                         "kernel": "inflate_kernel", "kernel_ms": round(inf_ms, 3), "finalize_ms": round(k_fin / args.steps, 3),
                         "algorithmic_bytes": comp_bytes + out_bytes},
            "cpu_baseline": cpu,
            "e2e": e2e,
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
            "device_ms_per_step": round(dev_ms, 3),
            "phase_ms": {"huff_tokens": round(k_phase[0], 3), "lz_resolve": round(k_phase[1], 3), "general_decoder": round(k_phase[2], 3),
                         "finalize": round(k_phase[3], 3), "fast_path_streams": list(ctx.last_fast_stats())},
            "checksums": checks,
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
