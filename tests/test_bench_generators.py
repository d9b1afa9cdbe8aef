"""CPU checks of the synthetic batches bench.py / tools/bench_mixed.py time on the GPU: the streams are what they claim to be
(the oracle - checker only - decodes them completely with the dictionary / container they name)."""
import numpy as np

import bench as B
from oracle import oracle as O


def _stream(m, i):
    return m["arena"][int(m["off"][i]):int(m["off"][i]) + int(m["ln"][i])].tobytes()


def test_dict_batch_streams_need_and_accept_their_dictionary():
    m = B.build_dict_batch(24, n_distinct=6)
    assert len(m["ln"]) == 24 and (m["mode"] == 0x81).all() and (m["off"] % 16 == 0).all()
    for i in range(6):
        s = _stream(m, i)
        assert s[1] & 0x20                                             # FDICT
        out, r = O.inflate_oneshot(s, dictionary=m["dictionary"], mode=O.MODE_INFLATER)
        assert r.success and r.out_len == 30000 and r.out_len <= int(m["cap"][i])
        _, r2 = O.inflate_oneshot(s, mode=O.MODE_INFLATER)
        assert r2.thrown_append == O.THROW_DICT_REQUIRED
    assert _stream(m, 7) == _stream(m, 1)                              # tiled


def test_mixed_batch_has_every_container_and_block_kind():
    m = B.build_mixed_batch(1024, 4, 0)
    n = len(m["ln"])
    assert n == 1024 and (m["off"] % 16 == 0).all()
    kinds = {"gzip": 0, "zlib": 0, "raw": 0, "dict": 0, "stored": 0}
    total = 0
    for i in range(n):
        s = _stream(m, i)
        hd = bool(m["mode"][i] & 0x80)
        out, r = O.inflate_oneshot(s, dictionary=m["dictionary"] if hd else None, mode=int(m["mode"][i] & 0x7f))
        assert r.thrown_append == 0 and r.out_len <= int(m["cap"][i]), i
        total += r.out_len
        kinds["dict"] += hd
        kinds["gzip"] += r.container == 2
        kinds["zlib"] += r.container == 1
        kinds["raw"] += r.container == 0
        kinds["stored"] += int(m["ln"][i]) >= r.out_len > 1000
    assert all(v > 0 for v in kinds.values()), kinds
    assert total == m["plain_bytes"]
