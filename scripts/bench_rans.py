"""Host rANS coder throughput (CPU only): synthetic discretised-Gaussian symbols over qarv's 64 scale tables.
usage: bench_rans.py [n_symbols] [n_streams] [threads]"""
import sys, time, ctypes as C
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT / 'lossy-vae_b200', ROOT / 'oracle'):
    sys.path.insert(0, str(p))
import os
from lvae import _native as N
if os.environ.get('LVAE_LIB_PATH'):
    N._LIB_PATH = Path(os.environ['LVAE_LIB_PATH'])
lib = N.lib()
import lvae_oracle as O
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
n_streams = int(sys.argv[2]) if len(sys.argv) > 2 else 16
threads = int(sys.argv[3]) if len(sys.argv) > 3 else 1
cdf, cdf_len, offset = [np.ascontiguousarray(np.asarray(t), dtype=np.int32) for t in O.build_cdf_tables()]
scales = np.asarray(O.default_scale_table())
rng = np.random.default_rng(0)
idx = rng.integers(0, cdf.shape[0], size=n).astype(np.int32)
sym = np.rint(rng.standard_normal(n) * scales[idx]).astype(np.int32)
begin = np.linspace(0, n, n_streams + 1).astype(np.int64)
cap = np.array([lib.lvae_rans_bound(int(begin[i + 1] - begin[i])) for i in range(n_streams)], dtype=np.int64)
out_begin = np.concatenate([[0], np.cumsum(cap)]).astype(np.int64)
out = np.empty(int(out_begin[-1]), dtype=np.uint8)
out_len = np.zeros(n_streams, dtype=np.int64)
p = lambda a: a.ctypes.data_as(C.c_void_p)
def enc():
    rc = lib.lvae_rans_encode_streams(p(sym), p(idx), p(begin), n_streams, p(cdf), cdf.shape[1], p(cdf_len), p(offset), cdf.shape[0],
                                      p(out), p(out_begin), p(out_len), threads)
    assert rc == 0, rc
enc()
reps = 9
def best(fn):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    return min(ts)
te = best(enc)
packed = np.concatenate([out[out_begin[i]:out_begin[i] + out_len[i]] for i in range(n_streams)])
in_begin = np.concatenate([[0], np.cumsum(out_len)]).astype(np.int64)
dec_out = np.empty(n, dtype=np.int32)
def dec():
    rc = lib.lvae_rans_decode_streams(p(packed), p(in_begin), p(idx), p(begin), n_streams, p(cdf), cdf.shape[1], p(cdf_len), p(offset),
                                      cdf.shape[0], p(dec_out), threads)
    assert rc == 0, rc
dec()
td = best(dec)
assert np.array_equal(dec_out, sym)
import hashlib
print(f'n={n} streams={n_streams} threads={threads}: encode {n / te / 1e6:.1f} Msym/s, decode {n / td / 1e6:.1f} Msym/s, '
      f'{packed.size * 8 / n:.3f} bits/sym, sha1 {hashlib.sha1(packed.tobytes()).hexdigest()[:12]}')
