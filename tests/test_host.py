"""Host-side logic (no GPU): C rANS coder + CDF builder against the oracle and the reference-made
fixtures, bit-stream container, model surface / state-dict contract, error conventions."""
import copy
import ctypes as C
import struct

import numpy as np
import pytest
import torch

import lvae_oracle as O


def _enc(lib, sym, idx, tables):
    cdf, clen, off = (np.ascontiguousarray(t.numpy()) for t in tables)
    sym = np.ascontiguousarray(sym, dtype=np.int32)
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    cap = int(lib.lvae_rans_bound(sym.size))
    buf = np.empty(cap, dtype=np.uint8)
    n = C.c_int64(0)
    rc = lib.lvae_rans_encode(sym.ctypes.data, idx.ctypes.data, sym.size, cdf.ctypes.data, cdf.shape[1],
                              clen.ctypes.data, off.ctypes.data, cdf.shape[0], buf.ctypes.data, cap, C.byref(n))
    assert rc == 0
    return buf[:n.value].tobytes()


def _dec(lib, data, idx, tables):
    cdf, clen, off = (np.ascontiguousarray(t.numpy()) for t in tables)
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    out = np.empty(idx.size, dtype=np.int32)
    d = np.frombuffer(data, dtype=np.uint8)
    rc = lib.lvae_rans_decode(d.ctypes.data, d.size, idx.ctypes.data, idx.size, cdf.ctypes.data, cdf.shape[1],
                              clen.ctypes.data, off.ctypes.data, cdf.shape[0], out.ctypes.data)
    return rc, out


@pytest.fixture(scope='module')
def tables():
    return O.build_cdf_tables()


def test_c_cdf_builder_equals_reference_tables(native_lib, golden):
    from lvae.models import entropy_coding as ec
    k = golden('entropy_kat')
    cdf, length, offset = ec.build_cdf_tables(ec.default_scale_table())
    assert np.array_equal(cdf.numpy(), k['cdf'])
    assert np.array_equal(length.numpy(), k['cdf_length'])
    assert np.array_equal(offset.numpy(), k['offset'])
    assert np.array_equal(ec.default_scale_table().numpy(), k['scale_table'])


def test_c_pmf_to_quantized_cdf_steals_for_zero_bins(native_lib):
    pmf = np.array([0.5, 1e-9, 0.25, 0.25 - 1e-9, 0.0], dtype=np.float32)
    out = np.zeros(6, dtype=np.int32)
    assert native_lib.lvae_pmf_to_quantized_cdf(pmf.ctypes.data, 5, 16, out.ctypes.data) == 0
    assert out.tolist() == O.pmf_to_quantized_cdf(pmf.tolist(), 16)
    assert out[-1] == 65536 and np.all(np.diff(out) >= 1)
    bad = np.array([0.5, float('nan')], dtype=np.float32)
    assert native_lib.lvae_pmf_to_quantized_cdf(bad.ctypes.data, 2, 16, out.ctypes.data) == -1


def test_c_rans_equals_python_oracle_and_roundtrips(native_lib, tables):
    g = torch.Generator().manual_seed(3)
    n = 6000
    idx = torch.randint(0, 64, (n,), generator=g).int().numpy()
    sym = torch.round(torch.randn(n, generator=g) * 4).int().numpy()
    sym[[5, 900, n - 1]] = [400, -300, 70000]           # bypass path: 4-bit nibbles
    data = _enc(native_lib, sym, idx, tables)
    assert data == O.rans_encode(sym.tolist(), idx.tolist(), *tables)
    rc, out = _dec(native_lib, data, idx, tables)
    assert rc == 0 and np.array_equal(out, sym)
    assert O.rans_decode(data, idx.tolist(), *tables) == sym.tolist()


def test_c_rans_golden_stream(native_lib, tables, golden):
    """The byte stream the reference produced for config 1 is reproduced from its symbols/indexes."""
    g = golden('qarv_rand_1x64x64')
    blob = g['bytes0'].tobytes()
    from lvae.utils import coding
    strings = coding.unpack_byte_string(blob[10:])
    assert len(strings) == 9
    for li, s in enumerate(strings):
        sym = g[f'sym{li}'][0].astype(np.int32).reshape(-1)
        idx = g[f'idx{li}'][0].astype(np.int32).reshape(-1)
        assert _enc(native_lib, sym, idx, tables) == s
        rc, out = _dec(native_lib, s, idx, tables)
        assert rc == 0 and np.array_equal(out, sym)


def test_c_rans_edge_cases(native_lib, tables):
    empty = np.zeros(0, dtype=np.int32)
    data = _enc(native_lib, empty, empty, tables)
    assert len(data) == 8                                 # just the flushed 64-bit state
    rc, out = _dec(native_lib, data, empty, tables)
    assert rc == 0 and out.size == 0
    sym = np.array([0, 1, -1, 2], dtype=np.int32)
    idx = np.array([0, 10, 63, 30], dtype=np.int32)
    data = _enc(native_lib, sym, idx, tables)
    rc, _ = _dec(native_lib, data[:4], idx, tables)       # truncated
    assert rc == -3
    rc, _ = _dec(native_lib, data, np.array([0, 10, 64, 30], dtype=np.int32), tables)   # index out of table
    assert rc == -1


def test_container_roundtrip_and_layout():
    from lvae.utils import coding
    strings = [b'', b'abc', bytes(range(256)) * 3, b'\x00']
    blob = coding.pack_byte_strings(strings)
    assert blob == O.pack_byte_strings(strings)
    assert blob[0] == 4 and struct.unpack('4I', blob[1:17]) == (0, 3, 768, 1)
    assert coding.unpack_byte_string(blob) == strings
    with pytest.raises(AssertionError):
        coding.unpack_byte_string(blob + b'x')


def test_bd_rate_identity_and_shift():
    from lvae.utils.coding import bd_rate
    r = [0.2, 0.4, 0.8, 1.6]
    p = [30.0, 33.0, 36.0, 39.0]
    assert abs(bd_rate(r, p, r, p)) < 1e-9
    assert abs(bd_rate(r, p, [x * 0.9 for x in r], p) + 10.0) < 1e-6


def test_model_surface_and_state_dict_contract(native_lib, sensitised_sd):
    import lvae
    torch.manual_seed(0)
    m = lvae.get_model('qarv_base')
    names = [k for k, _ in m.named_parameters()]
    want = dict(O.qarv_param_shapes())
    assert sorted(names) == sorted(want)
    for k, p in m.named_parameters():
        assert tuple(p.shape) == tuple(want[k]), k
    assert sum(p.numel() for p in m.parameters()) == 93_433_104 or round(sum(p.numel() for p in m.parameters()) / 1e6, 3) == 93.433
    sd = m.state_dict()
    assert len(sd) == 952                                   # 907 params + 45 buffers
    assert 'dec_blocks.0.discrete_gaussian._quantized_cdf' in sd
    assert 'dec_blocks.0.discrete_gaussian.likelihood_lower_bound.bound' in sd
    assert not any(k.endswith('scale_table') or k == '_dummy' for k in sd)
    missing, unexpected = m.load_state_dict(sensitised_sd, strict=False)
    assert not unexpected and all('discrete_gaussian' in k for k in missing)
    for attr in ('forward', 'forward_end2end', 'compress_mode', 'compress', 'decompress', 'compress_file',
                 'decompress_file', 'self_evaluate', 'conditional_sample', 'unconditional_sample', 'study',
                 'sample_lmb', 'expand_to_tensor'):
        assert callable(getattr(m, attr))
    assert m.num_latents == 9 and m.max_stride == 64 and m.lmb_range == (16.0, 2048.0) and m.default_lmb == 2048.0
    # optimizer param groups key on these substrings (lvae/trainer.py:182-194)
    assert all(('.weight' in k) or ('.bias' in k) or k.endswith('gamma') or k == 'bias' for k in names)
    m2 = copy.deepcopy(m)                                   # EMA (trainer.py:311)
    assert m2.__dict__['_engine'] is None and torch.equal(m2.bias, m.bias)
    str(m)


def test_default_init_matches_reference_conventions(native_lib):
    import lvae
    torch.manual_seed(0)
    m = lvae.get_model('qarv_base')
    sd = m.state_dict()
    assert float(sd['encoder.enc_blocks.1.gamma'].flatten()[0]) == pytest.approx(1e-6)
    assert torch.count_nonzero(sd['encoder.enc_blocks.0.bias']) == 0
    assert torch.count_nonzero(sd['dec_blocks.0.prior.bias']) == 0
    assert torch.count_nonzero(sd['bias']) == 0
    assert torch.count_nonzero(sd['encoder.enc_blocks.1.mlp.fc1.bias']) > 0      # Linear keeps torch default init


def test_no_cpu_fallback_and_error_conventions(native_lib):
    import lvae
    m = lvae.get_model('qarv_base').eval()
    im = torch.rand(1, 3, 64, 64)
    with pytest.raises(RuntimeError, match='no CPU path'):
        m(im, lmb=torch.tensor([64.0]))
    with pytest.raises(AssertionError):
        m(torch.rand(1, 3, 65, 64), lmb=torch.tensor([64.0]))
    with pytest.raises(ValueError):
        m.forward_end2end(im, 64.0, mode='bogus')
    with pytest.raises(AssertionError):
        m.compress(torch.rand(2, 3, 64, 64))
    with pytest.raises(KeyError):
        lvae.get_model('no_such_model')
    blk = m.dec_blocks[0]
    with pytest.raises(ValueError, match='Uninitialized CDFs'):
        blk.discrete_gaussian.host_tables()
    m.compress_mode()
    cdf, clen, off = blk.discrete_gaussian.host_tables()
    assert cdf.shape == (64, 249) and clen.max() == 249 and off.min() == -123


def test_qres_model_surface_and_state_dict_contract(native_lib, golden):
    """qres34m: parameter / buffer names of the reference's HierarchicalVAE, its default init, the
    CompressAI table set built by compress_mode(), and the no-CPU-path error."""
    import lvae
    import qres_oracle as Q
    torch.manual_seed(0)
    m = lvae.get_model('qres34m', lmb=64)
    want = dict(Q.qres_param_shapes())
    names = [k for k, _ in m.named_parameters()]
    assert sorted(names) == sorted(want)
    for k, p in m.named_parameters():
        assert tuple(p.shape) == tuple(want[k]), k
    sd = m.state_dict()
    assert len(sd) == 813 and round(sum(p.numel() for p in m.parameters()) / 1e6, 2) == 34.04
    for suffix in ('_offset', '_quantized_cdf', '_cdf_length', 'scale_table', 'scale_bound',
                   'likelihood_lower_bound.bound', 'lower_bound_scale.bound'):
        assert f'decoder.dec_blocks.0.discrete_gaussian.{suffix}' in sd
    assert sd['decoder.dec_blocks.0.discrete_gaussian.scale_table'].numel() == 0
    # default init conventions: zero-initialised prior head (zero_last), residual scaling of z_proj.2, gamma 1e-6
    assert torch.count_nonzero(sd['decoder.dec_blocks.0.prior.c4.weight']) == 0
    assert float(sd['encoder.enc_blocks.1.gamma'][0]) == pytest.approx(1e-6)
    assert m.out_net.mse_lmb == 64.0 and m.num_latents == 12 and m.max_stride == 64
    for attr in ('forward', 'forward_eval', 'forward_get_latents', 'uncond_sample', 'cond_sample', 'compress_mode',
                 'compress', 'decompress', 'compress_file', 'decompress_file'):
        assert callable(getattr(m, attr))
    with pytest.raises(RuntimeError, match='no CPU path'):
        m.eval()(torch.rand(1, 3, 64, 64))
    m.compress_mode()
    t = golden('qres_tables')
    dg = m.decoder.dec_blocks[3].discrete_gaussian
    assert np.array_equal(dg.scale_table.numpy(), t['scale_table'])
    cdf, clen, off = dg.host_tables()
    assert np.array_equal(cdf, t['cdf']) and np.array_equal(clen, t['cdf_length']) and np.array_equal(off, t['offset'])
    m2 = copy.deepcopy(m)
    assert m2.__dict__['_engine'] is None and torch.equal(m2.decoder.bias, m.decoder.bias)
    str(m)


def test_c_rans_encode_streams_equals_single_stream_calls(native_lib, tables):
    """the threaded multi-stream entry point produces, per stream, exactly the bytes of
    lvae_rans_encode, for ragged stream sizes (including an empty stream) and any thread count."""
    rng = np.random.default_rng(3)
    sizes = [5000, 1, 0, 777, 20000, 64]
    idx = rng.integers(0, 64, size=sum(sizes)).astype(np.int32)
    sym = np.rint(rng.normal(size=sum(sizes)) * (1 + idx * 0.3)).astype(np.int32)
    sym[::97] *= 40                                   # bypass-coded outliers
    cdf, clen, off = (np.ascontiguousarray(t.numpy()) for t in tables)
    begin = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    caps = [int(native_lib.lvae_rans_bound(n)) for n in sizes]
    out_begin = np.concatenate([[0], np.cumsum(caps)]).astype(np.int64)
    want = [_enc(native_lib, sym[begin[i]:begin[i + 1]], idx[begin[i]:begin[i + 1]], tables) for i in range(len(sizes))]
    for threads in (1, 3, 16):
        out = np.zeros(sum(caps), dtype=np.uint8)
        out_len = np.zeros(len(sizes), dtype=np.int64)
        rc = native_lib.lvae_rans_encode_streams(sym.ctypes.data, idx.ctypes.data, begin.ctypes.data, len(sizes),
                                                 cdf.ctypes.data, cdf.shape[1], clen.ctypes.data, off.ctypes.data, cdf.shape[0],
                                                 out.ctypes.data, out_begin.ctypes.data, out_len.ctypes.data, threads)
        assert rc == 0
        got = [out[out_begin[i]:out_begin[i] + out_len[i]].tobytes() for i in range(len(sizes))]
        assert got == want
    for i, n in enumerate(sizes):                     # and every stream decodes back
        rc, dec = _dec(native_lib, want[i], idx[begin[i]:begin[i + 1]], tables)
        assert rc == 0 and np.array_equal(dec, sym[begin[i]:begin[i + 1]])


def test_c_rans_all_symbols_out_of_table_grows_the_scratch(native_lib, tables):
    """Every symbol bypass-coded (40 bits each instead of <= 16): the encoder's scratch, sized for table symbols, must grow;
    bytes still equal the Python restatement."""
    rng = np.random.default_rng(11)
    n = 3000
    idx = rng.integers(0, 64, size=n).astype(np.int32)
    sym = (rng.integers(100000, 2000000, size=n) * rng.choice([-1, 1], size=n)).astype(np.int32)
    data = _enc(native_lib, sym, idx, tables)
    assert data == O.rans_encode(sym.tolist(), idx.tolist(), *tables)
    rc, out = _dec(native_lib, data, idx, tables)
    assert rc == 0 and np.array_equal(out, sym)


def test_c_rans_decode_streams_pairs_ragged_and_unaligned(native_lib, tables):
    """lvae_rans_decode_streams: streams are decoded two at a time in lockstep per worker; ragged sizes (incl. empty and a
    lone last stream), byte offsets that are not word-aligned, any thread count; a corrupt stream is reported."""
    rng = np.random.default_rng(5)
    sizes = [4000, 3, 0, 9001, 257, 1200, 77]
    idx = rng.integers(0, 64, size=sum(sizes)).astype(np.int32)
    sym = np.rint(rng.normal(size=sum(sizes)) * (1 + idx * 0.3)).astype(np.int32)
    sym[::53] *= 30
    cdf, clen, off = (np.ascontiguousarray(t.numpy()) for t in tables)
    begin = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    blobs = [_enc(native_lib, sym[begin[i]:begin[i + 1]], idx[begin[i]:begin[i + 1]], tables) for i in range(len(sizes))]
    for lead in (0, 1, 2):                            # leading pad bytes: misaligned stream starts
        packed = np.frombuffer(b'\x00' * lead + b''.join(blobs), dtype=np.uint8).copy()
        in_begin = (np.concatenate([[0], np.cumsum([len(b) for b in blobs])]) + lead).astype(np.int64)
        for threads in (1, 2, 3, 16):
            out = np.full(sum(sizes), -12345, dtype=np.int32)
            rc = native_lib.lvae_rans_decode_streams(packed.ctypes.data, in_begin.ctypes.data, idx.ctypes.data, begin.ctypes.data,
                                                     len(sizes), cdf.ctypes.data, cdf.shape[1], clen.ctypes.data, off.ctypes.data,
                                                     cdf.shape[0], out.ctypes.data, threads)
            assert rc == 0 and np.array_equal(out, sym), (lead, threads)
    trunc = np.frombuffer(blobs[3][:-8], dtype=np.uint8).copy()       # stream 3 without its last two words: it runs dry
    one_begin, sym_begin = np.array([0, trunc.size], dtype=np.int64), np.array([0, sizes[3]], dtype=np.int64)
    idx3 = idx[begin[3]:begin[4]].copy()
    out = np.zeros(sizes[3], dtype=np.int32)
    rc = native_lib.lvae_rans_decode_streams(trunc.ctypes.data, one_begin.ctypes.data, idx3.ctypes.data, sym_begin.ctypes.data, 1,
                                             cdf.ctypes.data, cdf.shape[1], clen.ctypes.data, off.ctypes.data, cdf.shape[0],
                                             out.ctypes.data, 1)
    assert rc == -3


@pytest.mark.parametrize('seed', range(6))
def test_c_rans_arbitrary_tables_equal_python_oracle(native_lib, seed):
    """Random (non-Gaussian) table sets: rows of length 1 .. 60 symbols with spiky, flat and long-tailed pmfs (many
    frequency-1 entries inside one decoder LUT bucket, single-symbol rows where everything but one value is bypass-coded),
    arbitrary offsets.  C encoder bytes == Python restatement, C decoder inverts them."""
    rng = np.random.default_rng(100 + seed)
    n_rows, stride = 12, 64
    cdf = np.zeros((n_rows, stride), dtype=np.int32)
    clen = np.zeros(n_rows, dtype=np.int32)
    off = rng.integers(-40, 5, size=n_rows).astype(np.int32)
    for r in range(n_rows):
        k = int(rng.integers(1, 61)) if r else 1                    # symbols incl. the escape symbol; row 0: escape only
        kind = r % 3
        pmf = (rng.random(k) ** 8 if kind == 0 else (np.ones(k) if kind == 1 else 1.0 / (1 + np.arange(k)) ** 3)).astype(np.float32)
        pmf = np.maximum(pmf / pmf.sum(), 0).astype(np.float32)
        out = np.zeros(k + 1, dtype=np.int32)
        assert native_lib.lvae_pmf_to_quantized_cdf(pmf.ctypes.data, k, 16, out.ctypes.data) == 0
        cdf[r, :k + 1] = out
        clen[r] = k + 1
    tables = (torch.from_numpy(cdf), torch.from_numpy(clen), torch.from_numpy(off))
    n = 4000
    idx = rng.integers(0, n_rows, size=n).astype(np.int32)
    sym = (off[idx] + rng.integers(-3, 66, size=n)).astype(np.int32)       # in-table and out-of-table on both sides
    sym[::401] = rng.integers(-10 ** 6, 10 ** 6, size=sym[::401].size)
    data = _enc(native_lib, sym, idx, tables)
    assert data == O.rans_encode(sym.tolist(), idx.tolist(), *tables)
    rc, dec = _dec(native_lib, data, idx, tables)
    assert rc == 0 and np.array_equal(dec, sym)
    assert O.rans_decode(data, idx.tolist(), *tables) == sym.tolist()


def test_c_rans_rejects_malformed_tables(native_lib, tables):
    cdf, clen, off = (np.ascontiguousarray(t.numpy()).copy() for t in tables)
    sym, idx = np.zeros(4, dtype=np.int32), np.zeros(4, dtype=np.int32)
    buf, n = np.empty(1024, dtype=np.uint8), C.c_int64(0)
    bad = cdf.copy()
    bad[0, 3] = bad[0, 2]                                           # not strictly increasing
    rc = native_lib.lvae_rans_encode(sym.ctypes.data, idx.ctypes.data, 4, bad.ctypes.data, bad.shape[1], clen.ctypes.data,
                                     off.ctypes.data, bad.shape[0], buf.ctypes.data, 1024, C.byref(n))
    assert rc == -1
    bad_len = clen.copy()
    bad_len[1] = cdf.shape[1] + 1                                    # longer than the row stride
    rc = native_lib.lvae_rans_encode(sym.ctypes.data, idx.ctypes.data, 4, cdf.ctypes.data, cdf.shape[1], bad_len.ctypes.data,
                                     off.ctypes.data, cdf.shape[0], buf.ctypes.data, 1024, C.byref(n))
    assert rc == -1


def test_overlay_mode_loads_trainer_from_the_reference_checkout():
    """LVAE_REFERENCE_ROOT: lvae.trainer / lvae.datasets / lvae.utils.general come, unmodified, from the reference checkout
    (with the timm / wandb they need -- here the test shims), bound to THIS package's models; without the variable nothing
    changes.  Skipped where the reference tree is absent."""
    import os
    import subprocess
    import sys
    import ref_loader
    if not ref_loader.available():
        pytest.skip('reference tree not present')
    root = os.path.join(os.path.dirname(__file__), '..')
    code = (
        "import sys, lvae, lvae.trainer, lvae.datasets, lvae.utils\n"
        "import lvae.models.qarv.model as qm\n"
        "assert lvae.trainer.__file__.startswith('/root/reference'), lvae.trainer.__file__\n"
        "assert lvae.datasets.__file__.startswith('/root/reference')\n"
        "assert not qm.__file__.startswith('/root/reference') and not lvae.evaluation.__file__.startswith('/root/reference')\n"
        "assert lvae.trainer.get_model is lvae.get_model and hasattr(lvae.utils, 'SimpleTable') and hasattr(lvae.utils.coding, 'bd_rate')\n"
        "assert hasattr(lvae.trainer.BaseTrainingWrapper, 'training_loop') or hasattr(lvae.trainer.BaseTrainingWrapper, 'main')\n"
        "print('overlay ok')\n")
    env = dict(os.environ, LVAE_REFERENCE_ROOT='/root/reference',
               PYTHONPATH=os.pathsep.join([os.path.join(root, 'lossy-vae_b200'), str(ref_loader.SHIMS)]))
    r = subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and 'overlay ok' in r.stdout, r.stderr[-2000:]
    env.pop('LVAE_REFERENCE_ROOT')
    r = subprocess.run([sys.executable, '-c', 'import lvae.trainer'], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and 'No module named' in r.stderr


def test_training_aten_restatements_equal_the_oracle_ops(native_lib):
    """lvae/training.py differentiates the head convolutions, VDBlocks and (in the cross-check mode) whole ConvNeXt blocks
    through ATen restatements on NHWC tensors; on CPU they must equal the oracle's NCHW ops value for value."""
    import torch.nn.functional as F
    import qres_oracle as Q
    from lvae import training as T
    g = torch.Generator().manual_seed(0)
    B, H, W, C_, k = 2, 6, 5, 8, 7
    x = torch.randn(B, C_, H, W, generator=g)
    nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous()
    # ConvNeXt block with AdaLN (common.py:142-161) and with affine LayerNorm (qresvae/model.py:163-182)
    sd = {'b.conv_dw.weight': torch.randn(C_, 1, k, k, generator=g) / k, 'b.conv_dw.bias': torch.randn(C_, generator=g),
          'b.embedding_layer.1.weight': torch.randn(2 * C_, 16, generator=g) / 4, 'b.embedding_layer.1.bias': torch.randn(2 * C_, generator=g),
          'b.mlp.fc1.weight': torch.randn(2 * C_, C_, generator=g) / 3, 'b.mlp.fc1.bias': torch.randn(2 * C_, generator=g),
          'b.mlp.fc2.weight': torch.randn(C_, 2 * C_, generator=g) / 4, 'b.mlp.fc2.bias': torch.randn(C_, generator=g),
          'b.gamma': torch.randn(1, C_, 1, 1, generator=g), 'b.norm.weight': torch.randn(C_, generator=g), 'b.norm.bias': torch.randn(C_, generator=g)}
    emb = torch.randn(B, 16, generator=g)
    ada = F.linear(F.gelu(emb), sd['b.embedding_layer.1.weight'], sd['b.embedding_layer.1.bias'])
    a = T._dwln_aten(nhwc(x), ada, sd['b.conv_dw.weight'], sd['b.conv_dw.bias'], None, None, k)
    y = F.linear(F.gelu(F.linear(a, sd['b.mlp.fc1.weight'], sd['b.mlp.fc1.bias'])), sd['b.mlp.fc2.weight'], sd['b.mlp.fc2.bias'])
    got = nhwc(x) + y * sd['b.gamma'].reshape(-1)
    assert torch.allclose(got, nhwc(O.convnext_block(sd, 'b.', x, emb)), atol=1e-5)
    sdq = dict(sd, **{'b.gamma': sd['b.gamma'].reshape(-1)})
    a = T._dwln_aten(nhwc(x), None, sd['b.conv_dw.weight'], sd['b.conv_dw.bias'], sd['b.norm.weight'], sd['b.norm.bias'], k)
    y = F.linear(F.gelu(F.linear(a, sd['b.mlp.fc1.weight'], sd['b.mlp.fc1.bias'])), sd['b.mlp.fc2.weight'], sd['b.mlp.fc2.bias'])
    assert torch.allclose(nhwc(x) + y * sdq['b.gamma'], nhwc(Q.convnext_block(sdq, 'b.', x)), atol=1e-5)
    # VDBlock on a channel concat (qresvae/model.py:143-149,270)
    x1 = torch.randn(B, 4, H, W, generator=g)
    vd = {f'v.c{i}.weight': w for i, w in enumerate([torch.randn(6, C_ + 4, 1, 1, generator=g), torch.randn(6, 6, 3, 3, generator=g) / 3,
                                                     torch.randn(6, 6, 3, 3, generator=g) / 3, torch.randn(5, 6, 1, 1, generator=g)], 1)}
    vd.update({f'v.c{i}.bias': torch.randn(n, generator=g) for i, n in zip(range(1, 5), (6, 6, 6, 5))})
    ps = [vd[f'v.c{i}.{t}'] for i in range(1, 5) for t in ('weight', 'bias')]
    assert torch.allclose(T._vd_aten(nhwc(x), nhwc(x1), *ps), nhwc(Q.vdblock(vd, 'v.', torch.cat([x, x1], 1))), atol=1e-5)
    # convolutions: 3x3 head, 1x1 + residual, patch down, 1x1 + pixel shuffle (NHWC and the final NCHW form)
    w3, b3 = torch.randn(5, C_, 3, 3, generator=g) / 3, torch.randn(5, generator=g)
    assert torch.allclose(T._conv_aten(nhwc(x), None, None, w3, b3, dict(stride=1, pad=1)), nhwc(F.conv2d(x, w3, b3, padding=1)), atol=1e-5)
    w1, b1 = torch.randn(C_, 4, 1, 1, generator=g), torch.randn(C_, generator=g)
    assert torch.allclose(T._conv_aten(nhwc(x1), None, nhwc(x), w1, b1, dict(stride=1, pad=0)), nhwc(x + F.conv2d(x1, w1, b1)), atol=1e-5)
    xe = torch.randn(B, C_, 8, 6, generator=g)
    wd, bd = torch.randn(12, C_, 2, 2, generator=g), torch.randn(12, generator=g)
    assert torch.allclose(T._conv_aten(nhwc(xe), None, None, wd, bd, dict(stride=2, pad=0)), nhwc(F.conv2d(xe, wd, bd, stride=2)), atol=1e-5)
    wu, bu = torch.randn(12, C_, 1, 1, generator=g), torch.randn(12, generator=g)
    up = F.pixel_shuffle(F.conv2d(x, wu, bu), 2)
    assert torch.allclose(T._conv_aten(nhwc(x), None, None, wu, bu, dict(stride=1, pad=0, r=2)), nhwc(up), atol=1e-5)
    assert torch.allclose(T._conv_aten(nhwc(x), None, None, wu, bu, dict(stride=1, pad=0, r=2, nchw_out=1)), up, atol=1e-5)


def test_gemm_desc_layout_matches_the_header_and_the_integration_stub(tmp_path):
    """the ctypes mirror of lvae_gemm_desc -- lvae/_native.py:GemmDesc AND the stub a maintainer would
    paste from INTEGRATION.md -- must have the size and field offsets gcc gives the struct in include/lvae_b200.h."""
    import ctypes as C
    import re
    import subprocess
    from pathlib import Path
    from lvae import _native as N
    ROOT = Path(__file__).resolve().parent.parent
    names = [f[0] for f in N.GemmDesc._fields_]
    src = tmp_path / 'layout.c'
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "lvae_b200.h"\nint main(void) {\n'
                   '  printf("%zu\\n", sizeof(lvae_gemm_desc));\n'
                   + ''.join(f'  printf("%zu\\n", offsetof(lvae_gemm_desc, {n}));\n' for n in names) + '  return 0;\n}\n')
    exe = tmp_path / 'layout'
    subprocess.run(['gcc', '-I', str(ROOT / 'include'), str(src), '-o', str(exe)], check=True)
    out = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert out[0] == C.sizeof(N.GemmDesc)
    assert out[1:] == [getattr(N.GemmDesc, n).offset for n in names]
    # the stub in INTEGRATION.md: same field names in the same order, same size
    md = (ROOT / 'INTEGRATION.md').read_text()
    block = md[md.index('class GemmDesc(C.Structure)'):]
    block = block[:block.index('def _check')]
    stub_names = re.findall(r"\('(\w+)',", block)
    assert stub_names == names, (stub_names, names)
    assert f'C.sizeof(GemmDesc) == {out[0]}' in block


def test_lossless_model_surface_state_dict_and_tables(native_lib, golden):
    """qres34m_lossless (reference qresvae/zoo.py:63-114, model.py:16-94): parameter names / shapes of the
    reference (incl. out_net.conv_mean / conv_scale), the seeded default init equal to the live reference's when it is
    present, and the 128-scale table set GaussianNLLOutputNet.update() builds."""
    import lvae
    import qres_oracle as Q
    import ref_loader
    torch.manual_seed(0)
    m = lvae.get_model('qres34m_lossless')
    want = dict(Q.qres_param_shapes(Q.qres34m_lossless_arch()))
    assert sorted(k for k, _ in m.named_parameters()) == sorted(want)
    for k, p in m.named_parameters():
        assert tuple(p.shape) == tuple(want[k]), k
    assert m.lossless and m.out_net.loss_name == 'nll' and m.num_latents == 12
    mine = {k: v.clone() for k, v in m.state_dict().items()}
    m.compress_mode()
    t = golden('qresll_tables')
    dg = m.out_net.discrete_gaussian
    assert np.array_equal(dg.scale_table.numpy(), t['scale_table'])
    cdf, clen, off = dg.host_tables()
    assert np.array_equal(cdf, t['cdf']) and np.array_equal(clen, t['cdf_length']) and np.array_equal(off, t['offset'])
    if ref_loader.available():
        ref = ref_loader.load_reference()
        try:
            torch.manual_seed(0)
            r = ref.get_model('qres34m_lossless')
            for k, v in r.state_dict().items():
                if 'discrete_gaussian' not in k:
                    assert torch.equal(v, mine[k]), k
        finally:
            ref_loader.unload_reference()


def test_gemm_tile_width_selection_is_a_pure_function_of_the_shape(native_lib):
    """lvae_gemm_tile_width (host-only): the N-tile width of the tensor-core GEMM.  Large GEMMs keep the widest tile that divides N
    (<= 128 with two operand planes, <= 256 with one); a 2-plane GEMM that would fill at most two waves of a 148-SM device gets
    the width that minimises waves x (K / 16) x (128 + 1.25 BN) -- narrower, so that more SMs work.  The GPU test
    test_gemm_tile_width_does_not_change_bits is what makes this choice free of consequences for the results."""
    from lvae import _native as NV
    tw = native_lib.lvae_gemm_tile_width
    f16x3, bf16 = NV.PREC_F16X3, NV.PREC_BF16
    for M, Nn, K in [(49152, 384, 768), (49152, 768, 384), (12288, 1024, 512), (196608, 384, 192)]:      # many waves: unchanged
        assert tw(M, Nn, K, f16x3, 148) == 128
    assert tw(196608, 192, 128, f16x3, 148) == 96 and tw(49152, 448, 256, f16x3, 148) == 112               # N / 2, N / 4
    assert tw(768, 512, 1024, f16x3, 148) == 32 and tw(768, 1024, 512, f16x3, 148) == 48                    # H/64 at batch 8
    assert tw(3072, 1024, 512, f16x3, 148) == 96 and tw(3072, 512, 1024, f16x3, 148) == 96                  # H/32
    assert tw(768, 512, 1024, bf16, 148) == 256 and tw(49152, 384, 768, bf16, 148) == 192                   # one plane: N only
    for M in (768, 1536, 3072, 6144):           # monotone: more rows never make the tile narrower
        assert tw(M, 512, 1024, f16x3, 148) <= tw(2 * M, 512, 1024, f16x3, 148)
    for bn in (tw(M, Nn, 512, f16x3, 148) for M in (128, 768, 5000) for Nn in (64, 200, 448, 1024)):
        assert bn % 16 == 0 and 16 <= bn <= 128
    assert tw(0, 512, 512, f16x3, 148) < 0 and tw(768, 512, 512, NV.PREC_FP32, 148) < 0
