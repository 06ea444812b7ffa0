"""QarvEngine: compiles a qarv model into a static sequence of liblvae_b200 kernel launches.

The reference runs the model as ~2000 ATen ops per forward with NCHW<->NHWC permutes around every
block (lvae/models/common.py:142-161, lvae/models/qarv/model.py:294-363).  Here a *plan* is built
once per (batch, height, width, mode): activations stay resident in HBM as NHWC fp32 matrices
[M = B*h*w, C]; every ConvNeXt block is three launches

    dwconv+LayerNorm+AdaLN  ->  fc1 GEMM + GELU  ->  fc2 GEMM * gamma + residual

the AdaLN projections of all 90 blocks are one launch, each latent layer's prior transform +
quantise + likelihood + per-image rate partial sums are one launch, and the loss assembly is one
launch.  The plan is replayed through a CUDA graph.  torch is used for device memory, streams and
graphs only.

Dense contractions run in one of three precisions (model.precision):
    'bf16x3'  tcgen05 tensor cores, operands split into bf16 (hi, lo) pairs, 3 MMAs per product,
              fp32 accumulation in TMEM -- the parity mode
    'bf16'    tcgen05, single bf16 pass -- fast, non-parity
    'fp32'    fp32 FFMA on CUDA cores
"""
import ctypes as C
import math
import time

import numpy as np
import torch

from . import _native as N
from .models import common


def _ptr(t):
    return 0 if t is None else t.data_ptr()


class _Op:
    __slots__ = ('fn', 'args', 'keep', 'name', 'meta')

    def __init__(self, name, fn, args, keep=None, meta=None):
        self.name, self.fn, self.args, self.keep, self.meta = name, fn, args, keep, meta or {}


class Plan:
    """A static launch list + its device buffers for one (B, H, W, mode)."""

    def __init__(self, engine, B, H, W, mode, want_elem):
        self.eng, self.B, self.H, self.W, self.mode, self.want_elem = engine, B, H, W, mode, want_elem
        self.dev = engine.device
        self.segments = [[]]      # op lists split at host interaction points
        self.enc_gelu = {}        # qres: planes of gelu(encoder feature) per resolution
        self.bufs = {}
        self.retired = []
        self.graphs = None
        self.n_launch = 0

    # ---- buffers
    def f32(self, *shape):
        return torch.empty(shape, dtype=torch.float32, device=self.dev)

    def i32(self, *shape):
        return torch.empty(shape, dtype=torch.int32, device=self.dev)

    def i16(self, *shape):
        return torch.empty(shape, dtype=torch.bfloat16, device=self.dev)      # raw 16-bit plane storage

    def named(self, name, numel, dtype=torch.float32):
        t = self.bufs.get(name)
        if t is None or t.numel() < numel:
            if t is not None:
                self.retired.append(t)     # earlier ops hold raw pointers into it
            t = torch.empty(numel, dtype=dtype, device=self.dev)
            self.bufs[name] = t
        return t

    # ---- op emission
    def op(self, name, fn, *args, keep=None, meta=None):
        """meta: dict(kind='gemm'|'dwln'|'latent'|'misc', flops=..., bytes=...) -- ALGORITHMIC work of the launch,
        used by bench.py for the roofline figures."""
        self.segments[-1].append(_Op(name, fn, args, keep, meta))
        self.n_launch += 1

    def cut(self):
        self.segments.append([])

    def run_segment(self, i, stream):
        for o in self.segments[i]:
            rc = o.fn(*o.args, stream)
            if rc != 0:
                N.check(rc, o.name)


class QarvEngine:
    def __init__(self, model):
        self.model = model
        # 'qarv' (AdaLN blocks, discretised-Gaussian latents) | 'rd' (continuous latents) | 'qres' (affine-LN blocks,
        # VDBlock heads, CompressAI's stock erfc likelihood, no lambda embedding)
        self.family = getattr(model, 'family', 'qarv')
        self.lib = N.lib()
        self.device = None
        self._wver = None
        self._epoch = 0                # bumped by invalidate()
        self._plans = {}               # key -> Plan, in least-recently-used order (see _get_plan)
        self.max_plans = int(__import__('os').environ.get('LVAE_MAX_PLANS', '6'))
        self.use_graphs = True
        # lvae_convnext_mlp for C <= 192 (False / LVAE_FUSE_MLP=0: the unfused GEMM pair, bit-identical)
        self.fuse_mlp = __import__('os').environ.get('LVAE_FUSE_MLP', '1') != '0'
        # producers write operand planes for prior / post_merge (False / LVAE_PLANE_CHAIN=0: fp32 + split pass)
        self.plane_chain = __import__('os').environ.get('LVAE_PLANE_CHAIN', '1') != '0'
        self.host_coder_s = 0.0        # seconds spent in the host rANS coder (bench.py --workload codec reads it)
        self.coder_threads = min(16, __import__('os').cpu_count() or 1)
        # batched decode in two half-batches, host rANS of one half under the GPU segment of the other.  Opt-in
        # (LVAE_DECODE_PINGPONG=1): at 8 images per call it is SLOWER (decompress 17.8 -> 23.4 ms) -- a layer's host time is the
        # latency of its longest single rANS stream, not a throughput: 8 streams on 16 threads take as long as 4, so two
        # half-batch calls double the host time (5 -> 8.5 ms) and the host, not the GPU, is what a layer waits for
        self.decode_pingpong = __import__('os').environ.get('LVAE_DECODE_PINGPONG', '0') == '1'
        # eval / compress plans of the qarv family: latent arithmetic as the epilogue of the posterior convolution (lvae_gemm_latent;
        # qm never reaches HBM, 9 launches fewer).  Opt-in (LVAE_FUSE_LATENT=1): same symbols, same rate, but SLOWER at batch 8 --
        # the posterior GEMMs are sub-wave launches (96 tiles at H/16: one tile per CTA, nothing to overlap the epilogue with) and a
        # tile has 8 epilogue warps for 128 x zdim elements of 336 instructions each: H/16 head 43.5 -> 117.5 us, where the
        # stand-alone kernel spreads the same arithmetic over every SM at full occupancy (13 us).  562 -> 556 images/s.
        self.fuse_latent = __import__('os').environ.get('LVAE_FUSE_LATENT', '0') == '1'
        self.blocks = [m for m in model.modules() if isinstance(m, (common.ConvNeXtBlockAdaLN, common.ConvNeXtBlockLN))]
        self.ada_off = {}
        off = 0
        for b in self.blocks:
            if isinstance(b, common.ConvNeXtBlockAdaLN):
                self.ada_off[id(b)] = off
                off += 2 * b.dim
        self.ada_total = off
        self.w = {}
        # load_state_dict() bumps the version counters already; the hook makes the re-pack explicit and also covers
        # state dicts loaded with assign=True
        model.register_load_state_dict_post_hook(lambda module, incompatible: self.invalidate())

    # ------------------------------------------------------------------ weights
    def _weights_version(self):
        """What the packed operand planes were made from: device, precision mode, and for EVERY parameter its storage
        address and autograd version counter (an optimizer step, load_state_dict(), .to(), `p.data = ...` all change one of
        them).  A write THROUGH `.data` (p.data.copy_(), dist.broadcast(p.data), ...) changes neither: callers that do
        that must call invalidate() -- lvae's own code paths (GraphedTrainStep, the model's load_state_dict / _apply
        hooks) do."""
        tables = tuple(b.discrete_gaussian.scale_table.numel() for b in self.model.dec_blocks
                       if hasattr(b, 'discrete_gaussian'))[:1]      # qres: the table appears with compress_mode()
        ver = ptr = 0
        for p in self.model.parameters():
            ver += p._version
            ptr = (ptr * 1000003 + p.data_ptr()) & 0xFFFFFFFFFFFFFFFF
        return (self.model._dummy.device, self.model.precision, ver, ptr, self._epoch, tables)

    def invalidate(self):
        """Force a re-pack of the weights (and a rebuild of the launch plans, which hold raw weight pointers) at the next
        use.  Needed after in-place writes through `.data`, which no version counter records."""
        self._epoch += 1

    def _dev_f32(self, t):
        return t.detach().to(self.device, torch.float32).contiguous()

    def _pack_gemm_weight(self, w2d, bias, prec=None):
        """w2d: [N, K] fp32 on device -> dict(w, bias, planes=[bf16 [N,K]] * npl); prec: precision mode the planes are
        for (default: the model's)"""
        ent = dict(w=w2d.contiguous(), bias=None if bias is None else self._dev_f32(bias),
                   N=w2d.shape[0], K=w2d.shape[1], planes=[])
        npl, pfmt = (self.npl, self.pfmt) if prec is None else (N.NUM_PLANES[prec], N.PLANE_FORMAT[prec])
        if npl:
            n = ent['w'].numel()
            pl = [torch.empty(n, dtype=torch.bfloat16, device=self.device) for _ in range(npl)]
            ptrs = [_ptr(t) for t in pl] + [0] * (3 - npl)
            f16 = pfmt == N.PLANES_F16       # fp16 planes carry w * 2^8 (include/lvae_b200.h LVAE_PREC_F16X3)
            N.check(self.lib.lvae_split_planes(_ptr(ent['w']), ptrs[0], ptrs[1], ptrs[2], n, pfmt,
                                               N.F16_WEIGHT_SCALE if f16 else 1.0, self._stream()), 'split_planes')
            ent['planes'] = pl
        return ent

    def _conv_weight(self, conv, pad_c=0):
        """pad_c: zero input channels appended (the operand is channel-padded by lvae_pad_channels)."""
        w = self._dev_f32(conv.weight)                    # [N, C, kh, kw]
        if pad_c:
            w = torch.nn.functional.pad(w, (0, 0, 0, 0, 0, pad_c))
        n, c, kh, kw = w.shape
        ent = self._pack_gemm_weight(w.permute(0, 2, 3, 1).reshape(n, kh * kw * c), conv.bias)
        ent['ks'] = kh
        return ent

    def _vd_weights(self, vd):
        return {k: self._conv_weight(getattr(vd, k)) for k in ('c1', 'c2', 'c3', 'c4')}

    def _pack_up(self, w, mod, dev):
        """patch_upsample = 1x1 conv + PixelShuffle(r): the packed weight is row-permuted so that the shuffle is a pure
        store pattern -- packed row (i*r+j)*Co + c  <-  reference row c*r*r + i*r + j  (common.py:33-38)"""
        conv, r = mod[0], mod.rate
        wt = self._dev_f32(conv.weight).reshape(conv.out_channels, conv.in_channels)
        co = conv.out_channels // (r * r)
        perm = torch.arange(conv.out_channels, device=dev).reshape(co, r * r).t().reshape(-1)
        w[id(mod)] = self._pack_gemm_weight(wt[perm], self._dev_f32(conv.bias)[perm])
        if id(mod) in self.tail_ids:
            w[(id(mod), 'tail')] = self._pack_gemm_weight(wt[perm], self._dev_f32(conv.bias)[perm], prec=self.tail_prec)

    @staticmethod
    def _z_pad(zdim, ks):
        """channels lvae_pad_channels appends to z before the z_proj conv: the GEMM needs C % 4 == 0 and K % 8 == 0"""
        zp = zdim
        while zp % 4 or (ks * ks * zp) % 8:
            zp += 1
        return zp - zdim

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def refresh_weights(self, force=False):
        m = self.model
        dev = m._dummy.device
        if dev.type != 'cuda':
            raise RuntimeError('lvae (B200 build) has no CPU path: move the model to a CUDA device first '
                               '(model.to("cuda")); the arithmetic lives in liblvae_b200.so')
        ver = self._weights_version()
        if not force and ver == self._wver:
            return
        base, _, tail = m.precision.partition('+')
        if base not in N.PRECISIONS or tail not in ('', 'tail1') or (tail and base != 'f16x3'):
            raise ValueError(f"unknown precision mode {m.precision!r}: one of {sorted(N.PRECISIONS)} or 'f16x3+tail1'")
        if self.device != dev or self.__dict__.get('prec') != N.PRECISIONS[base] or self.__dict__.get('tail_prec', None) != (N.PREC_F16 if tail else None):
            self._plans.clear()
        self.device = dev
        self.prec = N.PRECISIONS[base]
        # 'f16x3+tail1': the blocks and up-samplers AFTER CompresionStopFlag (qarv/zoo.py:78-88: 9 blocks + 2 patch
        # up-samplers, 9.3 % of the dense FLOPs) cannot change a symbol or the rate -- they only render the reconstruction --
        # so inference plans run them with ONE fp16 operand plane (1 MMA per product instead of 3, half the operand bytes).
        # Measured d PSNR against the reference stays inside the 0.01 dB budget (tests/test_gpu_model.py, profiles/r2_parity.md).
        self.tail_prec = N.PREC_F16 if tail else None
        self.tail_ids = set()
        if self.tail_prec is not None:
            after = False
            for mod in m.dec_blocks:
                if isinstance(mod, common.CompresionStopFlag):
                    after = True
                elif after:
                    self.tail_ids.add(id(mod))
        self.npl = N.NUM_PLANES[self.prec]        # 16-bit planes per tensor-core operand (0: fp32 CUDA-core path)
        self.pfmt = N.PLANE_FORMAT[self.prec]     # their element format (bf16 | fp16)
        w = {}
        with torch.cuda.device(dev):
            for b in self.blocks:
                C_, k = b.dim, b.kernel_size
                w[id(b)] = dict(
                    dw_w=self._dev_f32(b.conv_dw.weight).reshape(C_, k * k).t().contiguous(),
                    dw_b=self._dev_f32(b.conv_dw.bias),
                    fc1=self._pack_gemm_weight(self._dev_f32(b.mlp.fc1.weight), b.mlp.fc1.bias),
                    fc2=self._pack_gemm_weight(self._dev_f32(b.mlp.fc2.weight), b.mlp.fc2.bias),
                    gamma=self._dev_f32(b.gamma).reshape(C_),
                )
                if isinstance(b, common.ConvNeXtBlockLN):       # affine LayerNorm instead of AdaLN
                    w[id(b)]['ln_w'], w[id(b)]['ln_b'] = self._dev_f32(b.norm.weight), self._dev_f32(b.norm.bias)
                if id(b) in self.tail_ids:                      # single-plane copies for the inference plans' tail
                    w[id(b)]['fc1_tail'] = self._pack_gemm_weight(w[id(b)]['fc1']['w'], b.mlp.fc1.bias, prec=self.tail_prec)
                    w[id(b)]['fc2_tail'] = self._pack_gemm_weight(w[id(b)]['fc2']['w'], b.mlp.fc2.bias, prec=self.tail_prec)
            if self.ada_total:
                ada = [b for b in self.blocks if id(b) in self.ada_off]
                w['ada_w'] = torch.cat([self._dev_f32(b.embedding_layer[1].weight) for b in ada], 0).contiguous()
                w['ada_b'] = torch.cat([self._dev_f32(b.embedding_layer[1].bias) for b in ada], 0).contiguous()
                for j in (0, 2):
                    w[f'emb{j}_w'] = self._dev_f32(m.lmb_embedding[j].weight)
                    w[f'emb{j}_b'] = self._dev_f32(m.lmb_embedding[j].bias)
                fr = self.w.get('freqs')          # constant: kept across refreshes (no H2D copy inside a captured training step)
                w['freqs'] = fr if fr is not None and fr.device == dev else common.sinusoidal_frequencies(m.lmb_embed_dim[0], m._sin_period).to(dev)
            w['bias'] = self._dev_f32(m.bias).reshape(-1)
            mods = list(m.encoder.enc_blocks) + list(m.dec_blocks)
            for mod in mods:
                kind = getattr(mod, 'op_kind', None)
                if getattr(mod, 'downsapmle', None) is not None:        # rd: ConvNeXt block followed by a patch conv
                    w[id(mod.downsapmle)] = self._conv_weight(mod.downsapmle)
                if kind == 'down':
                    w[id(mod)] = self._conv_weight(mod)
                elif kind == 'up':
                    self._pack_up(w, mod, dev)
                elif getattr(mod, 'is_latent_block', False) and self.family == 'qres':
                    ks0 = mod.z_proj[0].kernel_size[0]
                    zpad = self._z_pad(mod.zdim, ks0)
                    tab = mod.discrete_gaussian.scale_table
                    w[id(mod)] = dict(posterior=self._vd_weights(mod.posterior), prior=self._vd_weights(mod.prior),
                                      z_proj0=self._conv_weight(mod.z_proj[0], pad_c=zpad), z_pad=zpad,
                                      z_proj2=self._conv_weight(mod.z_proj[2]),
                                      table=tab.detach().to(dev, torch.float32).contiguous() if tab.numel() else None)
                elif getattr(mod, 'is_latent_block', False):
                    w[id(mod)] = dict(post_merge=self._conv_weight(mod.post_merge),
                                      posterior=self._conv_weight(mod.posterior),
                                      z_proj=self._conv_weight(mod.z_proj),
                                      prior=self._conv_weight(mod.prior))
                    if self.family == 'qarv':
                        w[id(mod)]['table'] = mod.discrete_gaussian.scale_table.detach().to(dev, torch.float32).contiguous()
            on = getattr(m, 'out_net', None)
            if on is not None and hasattr(on, 'conv_mean'):        # GaussianNLLOutputNet (qres34m_lossless): two shuffle heads
                self._pack_up(w, on.conv_mean, dev)
                self._pack_up(w, on.conv_scale, dev)
                dg = getattr(on, 'discrete_gaussian', None)       # appears with compress_mode()
                w['out_table'] = (dg.scale_table.detach().to(dev, torch.float32).contiguous()
                                  if dg is not None and dg.scale_table.numel() else None)
        self.w = w
        self._wver = ver
        # plans hold raw weight pointers -> rebuild them
        self._plans.clear()

    # ------------------------------------------------------------------ plan construction helpers
    def _gemm(self, P, name, a0, geom, went, out, epi=N.EPI_BIAS, a1=None, C1=0, gamma=None, res=None, r=0,
              a_planes=None, out_planes=None, a_act=0, a1_planes=None, out_planes_act=0, prec=None, latent=None):
        """geom = (B, H, W, C0, ksize, stride, pad) of the NHWC input a0.  In a tensor-core mode the A operand is
        either `a_planes` (bf16 planes written by the producing kernel) or a0/a1, which the library im2col-splits
        into the plan's workspace first."""
        B, H, W, C0, ks, st, pad = geom
        d = N.GemmDesc()
        d.a0, d.a1 = _ptr(a0), _ptr(a1)
        d.B, d.H, d.W, d.C0, d.C1 = B, H, W, C0, C1
        d.ksize, d.stride, d.pad = ks, st, pad
        d.w, d.bias, d.N = _ptr(went['w']), _ptr(went['bias']), went['N']
        d.epilogue, d.gamma, d.res, d.out = epi, _ptr(gamma), _ptr(res), _ptr(out)
        prec = self.prec if prec is None else prec
        npl = N.NUM_PLANES[prec]
        d.shuffle_r, d.precision, d.a_act, d.out_planes_act = r, prec, a_act, out_planes_act
        assert went['K'] == ks * ks * C0 + C1, (name, went['K'], ks, C0, C1)
        Mo = B * ((H + 2 * pad - ks) // st + 1) * ((W + 2 * pad - ks) // st + 1)
        ws = None
        if npl:
            N.set_planes(d, 'w', went['planes'])
            N.set_planes(d, 'a', a_planes)
            N.set_planes(d, 'out', out_planes)
            N.set_planes(d, 'a1', a1_planes)
            if a1_planes is not None:
                d.C1 = C1
            if a_planes is None:
                ws = P.named('tc_ws', Mo * went['K'] * npl, dtype=torch.bfloat16)
                d.workspace, d.workspace_bytes = _ptr(ws), ws.numel() * 2
        else:
            assert a_planes is None and out_planes is None
        meta = dict(kind='gemm', flops=2 * Mo * went['N'] * went['K'], M=Mo, N=went['N'], K=went['K'], terms=N.MMA_TERMS[prec],
                    bytes=4 * (B * H * W * (C0 + C1) + went['N'] * went['K'] + Mo * went['N'] * (2 if res is not None else 1)))
        if latent is not None:      # the implicit 3x3 posterior convolution with the latent arithmetic as its epilogue
            meta['latent_elems'] = Mo * went['N']
            P.op(name, self.lib.lvae_gemm_latent, C.byref(d), C.byref(latent),
                 keep=(d, latent, went, out, a_planes), meta=meta)
            return
        P.op(name, self.lib.lvae_gemm, C.byref(d), keep=(d, a0, a1, went, out, gamma, res, a_planes, out_planes, ws, a1_planes),
             meta=meta)

    def _block(self, P, blk, x, B, Hs, Ws, out=None, out_planes=None, planes_only=False, planes_act=0):
        """x: [M, C] fp32 NHWC; returns the output buffer (x itself when out is None: in place).  out_planes: bf16
        planes of the result, written by the fc2 epilogue for a tensor-core consumer (the 3x3 posterior conv)."""
        wb = self.w[id(blk)]
        C_, hid, k = blk.dim, blk.hidden, blk.kernel_size
        M = B * Hs * Ws
        out = x if out is None else out
        ln_w, ln_b = _ptr(wb.get('ln_w')), _ptr(wb.get('ln_b'))        # affine LayerNorm (qres) | 0: AdaLN
        ada_off = self.ada_off.get(id(blk), 0)
        # algorithmic bytes: read x fp32, write the GEMM operand (fp32, or npl bf16 planes)
        dw_meta = dict(kind='dwln', bytes=M * C_ * (4 + (2 * self.npl if self.npl else 4)), flops=2 * M * C_ * k * k)
        if self._tail(P, blk) and out_planes is None:
            # after the stop flag (inference plans of the 'f16x3+tail1' mode): one fp16 plane per operand, unfused GEMM pair
            tp = self.tail_prec
            A = [P.named('scratch_a0', M * C_, dtype=torch.bfloat16)]
            Hd = [P.named('scratch_h0', M * hid, dtype=torch.bfloat16)]
            dw_meta['bytes'] = M * C_ * (4 + 2)
            P.op('dwln', self.lib.lvae_dwconv_ln_adaln_planes, _ptr(x), _ptr(wb['dw_w']), _ptr(wb['dw_b']),
                 _ptr(P.ada), self.ada_total, ada_off, ln_w, ln_b, _ptr(A[0]), 0, 0, N.PLANE_FORMAT[tp], B, Hs, Ws, C_, k,
                 keep=(x, A), meta=dw_meta)
            self._gemm(P, 'fc1', None, (1, 1, M, C_, 1, 1, 0), wb['fc1_tail'], None, epi=N.EPI_BIAS_GELU, a_planes=A, out_planes=Hd, prec=tp)
            self._gemm(P, 'fc2', None, (1, 1, M, hid, 1, 1, 0), wb['fc2_tail'], out, epi=N.EPI_SCALE_RES,
                       gamma=wb['gamma'], res=x, a_planes=Hd, prec=tp)
            return out
        if self.npl:
            # tensor-core modes: the A operand of each GEMM travels as bf16 planes written by its producer
            A = [P.named(f'scratch_a{i}', M * C_, dtype=torch.bfloat16) for i in range(self.npl)]
            Hd = [P.named(f'scratch_h{i}', M * hid, dtype=torch.bfloat16) for i in range(self.npl)]
            ap = [_ptr(t) for t in A] + [0] * (3 - self.npl)
            P.op('dwln', self.lib.lvae_dwconv_ln_adaln_planes, _ptr(x), _ptr(wb['dw_w']), _ptr(wb['dw_b']),
                 _ptr(P.ada), self.ada_total, ada_off, ln_w, ln_b, ap[0], ap[1], ap[2], self.pfmt, B, Hs, Ws, C_, k,
                 keep=(x, A), meta=dw_meta)
            if self._mlp_fused(blk) and not planes_only:
                # narrow layers (H/4 stages): fc1 -> GELU -> fc2 -> layer scale + residual in one kernel, the hidden
                # tensor never leaves the SM (bit-identical to the two GEMMs below)
                w1, w2 = wb['fc1'], wb['fc2']
                op0, op1 = (_ptr(out_planes[0]), _ptr(out_planes[1])) if out_planes is not None else (0, 0)
                P.op('mlp', self.lib.lvae_convnext_mlp_planes, ap[0], ap[1], _ptr(w1['planes'][0]), _ptr(w1['planes'][1]),
                     _ptr(w1['bias']), _ptr(w2['planes'][0]), _ptr(w2['planes'][1]), _ptr(w2['bias']), _ptr(wb['gamma']),
                     _ptr(x), _ptr(out), op0, op1, planes_act, M, C_, hid, self.prec, keep=(x, out, A, wb, out_planes),
                     meta=dict(kind='gemm', flops=4 * M * C_ * hid, M=M, N=C_, K=hid, bytes=M * C_ * (4 + 4 + 4), terms=N.MMA_TERMS[self.prec]))
                return out
            self._gemm(P, 'fc1', None, (1, 1, M, C_, 1, 1, 0), wb['fc1'], None, epi=N.EPI_BIAS_GELU, a_planes=A, out_planes=Hd)
            # planes_only: the consumer reads the 16-bit planes, the fp32 result is never written
            self._gemm(P, 'fc2', None, (1, 1, M, hid, 1, 1, 0), wb['fc2'], None if planes_only else out, epi=N.EPI_SCALE_RES,
                       gamma=wb['gamma'], res=x, a_planes=Hd, out_planes=out_planes, out_planes_act=planes_act)
            return out
        A = P.named('scratch_a', M * C_)
        Hd = P.named('scratch_h', M * hid)
        P.op('dwln', self.lib.lvae_dwconv_ln_adaln, _ptr(x), _ptr(wb['dw_w']), _ptr(wb['dw_b']),
             _ptr(P.ada), self.ada_total, ada_off, ln_w, ln_b, _ptr(A), B, Hs, Ws, C_, k,
             keep=(x, A), meta=dw_meta)
        self._gemm(P, 'fc1', A, (1, 1, M, C_, 1, 1, 0), wb['fc1'], Hd, epi=N.EPI_BIAS_GELU)
        self._gemm(P, 'fc2', Hd, (1, 1, M, hid, 1, 1, 0), wb['fc2'], out, epi=N.EPI_SCALE_RES,
                   gamma=wb['gamma'], res=x)
        return out

    def _tail(self, P, mod):
        """True when `mod` lies after the stop flag and plan P may run it in the reduced tail precision (never a training
        plan: gradients and the train-mode loss keep the full operand split)."""
        return self.__dict__.get('tail_prec') is not None and id(mod) in self.tail_ids and P.mode != 'train'

    def _mlp_fused(self, blk):
        return (self.fuse_mlp and self.npl == 2 and blk.dim % 64 == 0 and blk.dim <= 192 and blk.hidden % 32 == 0)

    def _embedding(self, P, lmb):
        m, w, B = self.model, self.w, P.B
        P.ada = None
        if not self.ada_total:          # qres: no lambda embedding
            return
        E0, E1 = m.lmb_embed_dim
        emb0, e1, emb = P.f32(B, E0), P.f32(B, E1), P.f32(B, E1)
        P.ada = P.f32(B, self.ada_total)
        P.op('sinusoid', self.lib.lvae_lmb_sinusoid, _ptr(lmb), _ptr(w['freqs']), _ptr(emb0), B, E0,
             float(m._sin_period), float(m.MAX_LMB), keep=(lmb, emb0))
        P.op('emb0', self.lib.lvae_small_linear, _ptr(emb0), _ptr(w['emb0_w']), _ptr(w['emb0_b']), _ptr(e1),
             B, E0, E1, 0, 1, keep=(e1,))
        P.op('emb2', self.lib.lvae_small_linear, _ptr(e1), _ptr(w['emb2_w']), _ptr(w['emb2_b']), _ptr(emb),
             B, E1, E1, 0, 0, keep=(emb,))
        P.op('adaln_all', self.lib.lvae_small_linear, _ptr(emb), _ptr(w['ada_w']), _ptr(w['ada_b']), _ptr(P.ada),
             B, E1, self.ada_total, 1, 0)

    def _encoder(self, P, im):
        m, B, H, W = self.model, P.B, P.H, P.W
        feats = {}
        x, Cc, s = None, 3, 1
        pinned = False
        for mod in m.encoder.enc_blocks:
            kind = getattr(mod, 'op_kind', None)
            if kind == 'down':
                r = mod.rate
                s *= r
                Hs, Ws = H // s, W // s
                out = P.f32(B * Hs * Ws, mod.out_channels)
                if x is None:
                    patches = P.f32(B * Hs * Ws, Cc * r * r)
                    P.op('im2patch', self.lib.lvae_image_to_patches, _ptr(im), _ptr(patches), B, H, W, r,
                         float(m.im_shift), float(m.im_scale), keep=(im, patches))
                    self._gemm(P, 'down0', patches, (1, 1, B * Hs * Ws, Cc * r * r, 1, 1, 0), self.w[id(mod)], out)
                else:
                    self._gemm(P, 'down', x, (B, Hs * r, Ws * r, Cc, r, r, 0), self.w[id(mod)], out)
                x, Cc, pinned = out, mod.out_channels, False
            elif getattr(mod, 'downsapmle', None) is not None:
                # rd ConvNeXtAdaLNPatchDown (rd/model.py:16-24): the block's own output is not a tapped feature, the
                # tapped one (last plain block at this resolution) must survive -> block out of place, then patch conv
                Hs, Ws = H // s, W // s
                y = self._block(P, mod, x, B, Hs, Ws, out=P.named('enc_tmp', B * Hs * Ws * Cc)[:B * Hs * Ws * Cc])
                conv = mod.downsapmle
                r = conv.rate
                s *= r
                out = P.f32(B * (Hs // r) * (Ws // r), conv.out_channels)
                self._gemm(P, 'down', y, (B, Hs, Ws, Cc, r, r, 0), self.w[id(conv)], out)
                x, Cc, pinned = out, conv.out_channels, False
            elif isinstance(mod, (common.ConvNeXtBlockAdaLN, common.ConvNeXtBlockLN)):
                Hs, Ws = H // s, W // s
                x = self._block(P, mod, x, B, Hs, Ws, out=P.f32(B * Hs * Ws, Cc) if pinned else None)
                pinned = False
            elif isinstance(mod, common.SetKey):
                feats[mod.key] = x
                pinned = True
            else:
                raise TypeError(f'unsupported encoder module {type(mod)}')
            if self.family in ('rd', 'qres'):
                feats[H // s] = x          # keyed by feature height, last writer wins (rd/model.py:236-244, qresvae/model.py:200-206)
        return feats

    def _top_down(self, P, feats, nH, nW, latent_fn, stop_at_flag=False):
        """Walks dec_blocks.  latent_fn(P, blk, li, feature, prior, geom) emits the ops that produce z
        [M, zdim] for latent layer li and returns the z buffer."""
        m, B = self.model, P.B
        Hs, Ws = nH, nW
        Cc = m.dec_blocks[0].in_channels
        x = P.f32(B * Hs * Ws, Cc)
        P.op('bias', self.lib.lvae_broadcast_bias, _ptr(self.w['bias']), _ptr(x), B * Hs * Ws, Cc, keep=(x,))
        li = 0
        for mod in m.dec_blocks:
            kind = getattr(mod, 'op_kind', None)
            M = B * Hs * Ws
            if getattr(mod, 'is_latent_block', False):
                wl = self.w[id(mod)]
                zd = mod.zdim
                prior = P.f32(M, 2 * zd)
                if self.family == 'qres':
                    P.cur_xg = None
                    if self.npl == 2 and self.plane_chain and Cc % 64 == 0:
                        # both VDBlock heads start with gelu(feature): resnet_front's epilogue writes those planes once
                        P.cur_xg = [P.named(f'xg_pl{i}', M * Cc, dtype=torch.bfloat16)[:M * Cc] for i in range(self.npl)]
                    x = self._block(P, mod.resnet_front, x, B, Hs, Ws, out_planes=P.cur_xg, planes_act=1 if P.cur_xg is not None else 0)
                    self._vdblock(P, 'prior', wl['prior'], x, (B, Hs, Ws, Cc), prior, a_planes=P.cur_xg)
                elif self.npl and self.plane_chain and not self._mlp_fused(mod.resnet_front):
                    # resnet_front's fc2 epilogue also writes its result as planes: the prior head reads them directly
                    xp = [P.named(f'x_pl{i}', M * Cc, dtype=torch.bfloat16)[:M * Cc] for i in range(self.npl)]
                    x = self._block(P, mod.resnet_front, x, B, Hs, Ws, out_planes=xp)
                    self._gemm(P, 'prior', None, (B, Hs, Ws, Cc, 1, 1, 0), wl['prior'], prior, a_planes=xp)
                else:
                    x = self._block(P, mod.resnet_front, x, B, Hs, Ws)
                    self._gemm(P, 'prior', x, (B, Hs, Ws, Cc, 1, 1, 0), wl['prior'], prior)
                z = latent_fn(P, mod, li, x, prior, (B, Hs, Ws, Cc))
                li += 1
                if self.family == 'qres':
                    # z_proj = conv (3x3 | 1x1) -> GELU -> 1x1, added to the feature (qresvae/model.py:236-240,278)
                    zp, ks0 = wl['z_pad'], wl['z_proj0']['ks']
                    if zp:
                        zz = P.named('z_padded', M * (zd + zp))[:M * (zd + zp)]
                        P.op('pad_z', self.lib.lvae_pad_channels, _ptr(z), _ptr(zz), M, zd, zd + zp, keep=(z, zz))
                    else:
                        zz = z
                    hz = wl['z_proj0']['N']
                    t = P.named('z_hidden', M * hz)[:M * hz]
                    self._gemm(P, 'z_proj0', zz, (B, Hs, Ws, zd + zp, ks0, 1, (ks0 - 1) // 2), wl['z_proj0'], t, epi=N.EPI_BIAS_GELU)
                    self._gemm(P, 'z_proj2', t, (B, Hs, Ws, hz, 1, 1, 0), wl['z_proj2'], x, epi=N.EPI_BIAS_RES, res=x)
                else:
                    self._gemm(P, 'z_proj', z, (B, Hs, Ws, zd, 1, 1, 0), wl['z_proj'], x, epi=N.EPI_BIAS_RES, res=x)
                x = self._block(P, mod.resnet_end, x, B, Hs, Ws)
            elif isinstance(mod, (common.ConvNeXtBlockAdaLN, common.ConvNeXtBlockLN)):
                x = self._block(P, mod, x, B, Hs, Ws)
            elif kind == 'up':
                r = mod.rate
                co = mod[0].out_channels // (r * r)
                last = co == 3
                out = P.f32(B, 3, Hs * r, Ws * r) if last else P.f32(M * r * r, co)
                tail = self._tail(P, mod)
                self._gemm(P, 'up', x, (B, Hs, Ws, Cc, 1, 1, 0), self.w[(id(mod), 'tail')] if tail else self.w[id(mod)], out,
                           epi=N.EPI_SHUFFLE_NCHW if last else N.EPI_SHUFFLE_NHWC, r=r, prec=self.tail_prec if tail else None)
                x, Cc, Hs, Ws = out, co, Hs * r, Ws * r
            elif isinstance(mod, common.CompresionStopFlag):
                if stop_at_flag:
                    return None
            else:
                raise TypeError(f'unsupported decoder module {type(mod)}')
        return x

    def _vdblock(self, P, name, wv, a0, geom, out, a1=None, C1=0, a_planes=None, a1_planes=None):
        """VDBlock without residual (qresvae/model.py:143-149): c4(gelu(c3(gelu(c2(gelu(c1(gelu(x)))))))).  The first
        GELU is applied to the operand as it is read (a_act), the others ride in the producing GEMM's epilogue."""
        B, Hs, Ws, C0 = geom
        M = B * Hs * Ws
        hid, ks = wv['c1']['N'], wv['c2']['ks']
        if self.npl and ks == 3 and hid % 16 == 0 and self.plane_chain:
            # tensor-core modes: the hidden maps travel as 16-bit planes from epilogue to epilogue and the two 3x3 convs
            # read them implicitly (shifted TMA boxes) -- no fp32 round trip, no 9x im2col workspace
            h1 = [P.named(f'vd_h1_pl{i}', M * hid, dtype=torch.bfloat16)[:M * hid] for i in range(self.npl)]
            h2 = [P.named(f'vd_h2_pl{i}', M * hid, dtype=torch.bfloat16)[:M * hid] for i in range(self.npl)]
            if a_planes is not None:      # gelu(input) already travels as planes (written by the producers)
                self._gemm(P, name + '.c1', None, (B, Hs, Ws, C0, 1, 1, 0), wv['c1'], None, epi=N.EPI_BIAS_GELU, C1=C1,
                           a_planes=a_planes, a1_planes=a1_planes, out_planes=h1)
            else:
                self._gemm(P, name + '.c1', a0, (B, Hs, Ws, C0, 1, 1, 0), wv['c1'], None, epi=N.EPI_BIAS_GELU, a1=a1, C1=C1, a_act=1,
                           out_planes=h1)
            self._gemm(P, name + '.c2', None, (B, Hs, Ws, hid, 3, 1, 1), wv['c2'], None, epi=N.EPI_BIAS_GELU, a_planes=h1, out_planes=h2)
            self._gemm(P, name + '.c3', None, (B, Hs, Ws, hid, 3, 1, 1), wv['c3'], None, epi=N.EPI_BIAS_GELU, a_planes=h2, out_planes=h1)
            self._gemm(P, name + '.c4', None, (B, Hs, Ws, hid, 1, 1, 0), wv['c4'], out, a_planes=h1)
            return
        h1, h2 = P.named('vd_h1', M * hid)[:M * hid], P.named('vd_h2', M * hid)[:M * hid]
        self._gemm(P, name + '.c1', a0, (B, Hs, Ws, C0, 1, 1, 0), wv['c1'], h1, epi=N.EPI_BIAS_GELU, a1=a1, C1=C1, a_act=1)
        self._gemm(P, name + '.c2', h1, (B, Hs, Ws, hid, ks, 1, (ks - 1) // 2), wv['c2'], h2, epi=N.EPI_BIAS_GELU)
        self._gemm(P, name + '.c3', h2, (B, Hs, Ws, hid, ks, 1, (ks - 1) // 2), wv['c3'], h1, epi=N.EPI_BIAS_GELU)
        self._gemm(P, name + '.c4', h1, (B, Hs, Ws, hid, 1, 1, 0), wv['c4'], out)

    def _posterior(self, P, blk, x, enc_feat, geom, fuse=None):
        """transform_posterior (qarv/model.py:56-70) -> qm [M, zdim].  fuse = dict(lat=LatentEpilogue, z=buffer): when the head
        runs as the implicit tensor-core convolution, the latent arithmetic becomes its epilogue (lvae_gemm_latent): z is
        written instead of qm, fuse['done'] is set and None comes back."""
        B, Hs, Ws, Cc = geom
        M = B * Hs * Ws
        wl = self.w[id(blk)]
        if self.family == 'qres':        # posterior(cat([feature, enc_feature])) (qresvae/model.py:270)
            qm = P.f32(M, blk.zdim)
            eg = None
            if getattr(P, 'cur_xg', None) is not None and blk.enc_width % 8 == 0:
                # gelu(encoder feature) as planes, once per resolution (the latent blocks of a stage share the feature)
                eg = P.enc_gelu.get(Hs)
                if eg is None:
                    eg = [P.i16(M * blk.enc_width) for _ in range(self.npl)]
                    ptrs = [_ptr(t) for t in eg] + [0] * (3 - self.npl)
                    P.op('enc_gelu', self.lib.lvae_gelu_split_planes, _ptr(enc_feat), ptrs[0], ptrs[1], ptrs[2],
                         M * blk.enc_width, self.pfmt, keep=(enc_feat, eg))
                    P.enc_gelu[Hs] = eg
            self._vdblock(P, 'posterior', wl['posterior'], x, geom, qm, a1=enc_feat, C1=blk.enc_width,
                          a_planes=P.cur_xg if eg is not None else None, a1_planes=eg)
            return qm
        We = blk.enc_width
        mg = P.named('post_m', M * Cc)[:M * Cc]
        if self.npl and Cc % 64 == 0 and We % 8 == 0 and self.plane_chain:
            # the two branch blocks hand their results to post_merge as 16-bit planes (written by their fc2
            # epilogues, never as fp32): the K-concat reads both plane sets, no im2col / split pass
            pdt = torch.bfloat16
            ep = [P.named(f'post_e_pl{i}', M * We, dtype=pdt)[:M * We] for i in range(self.npl)]
            fp = [P.named(f'post_f_pl{i}', M * Cc, dtype=pdt)[:M * Cc] for i in range(self.npl)]
            self._block(P, blk.posterior0, enc_feat, B, Hs, Ws, out=mg, out_planes=ep, planes_only=True)
            self._block(P, blk.posterior1, x, B, Hs, Ws, out=mg, out_planes=fp, planes_only=True)
            self._gemm(P, 'post_merge', None, (B, Hs, Ws, Cc, 1, 1, 0), wl['post_merge'], mg, a_planes=fp, a1_planes=ep, C1=We)
        else:
            e = self._block(P, blk.posterior0, enc_feat, B, Hs, Ws, out=P.named('post_e', M * We)[:M * We])
            f = self._block(P, blk.posterior1, x, B, Hs, Ws, out=P.named('post_f', M * Cc)[:M * Cc])
            self._gemm(P, 'post_merge', f, (B, Hs, Ws, Cc, 1, 1, 0), wl['post_merge'], mg, a1=e, C1=We)
        qm = P.f32(M, wl['posterior']['N'])          # qarv: zdim means; rd: (mean_raw | std_raw) = 2 * zdim
        if self.npl and Cc % 64 == 0:
            # tensor-core modes: posterior2's fc2 epilogue also writes its result as bf16 planes, which the 3x3
            # head convolves implicitly (shifted TMA boxes) -- no im2col workspace
            mp = [P.named(f'post_m_pl{i}', M * Cc, dtype=torch.bfloat16)[:M * Cc] for i in range(self.npl)]
            mg = self._block(P, blk.posterior2, mg, B, Hs, Ws, out_planes=mp)
            if fuse is not None and wl['posterior']['N'] <= 128:
                self._gemm(P, 'posterior', None, (B, Hs, Ws, Cc, 3, 1, 1), wl['posterior'], fuse['z'], a_planes=mp, latent=fuse['lat'])
                fuse['done'] = True
                return None
            self._gemm(P, 'posterior', None, (B, Hs, Ws, Cc, 3, 1, 1), wl['posterior'], qm, a_planes=mp)
        else:
            mg = self._block(P, blk.posterior2, mg, B, Hs, Ws)
            self._gemm(P, 'posterior', mg, (B, Hs, Ws, Cc, 3, 1, 1), wl['posterior'], qm)
        return qm

    # ------------------------------------------------------------------ plans
    def _latent_layout(self, B, nH, nW):
        """per latent layer: (hw, zdim, n_partials, column offset)"""
        lay, off = [], 0
        Hs, Ws = nH, nW
        for mod in self.model.dec_blocks:
            if getattr(mod, 'is_latent_block', False):
                hw = Hs * Ws
                np_ = self.lib.lvae_latent_num_partials(hw, mod.zdim)
                if self.family == 'qarv':       # the latent epilogue of the posterior convolution has its own slot layout
                    np_ = max(np_, self.lib.lvae_gemm_latent_num_partials(Hs, Ws, mod.zdim))
                lay.append((hw, mod.zdim, np_, off, Hs, Ws))
                off += np_
            elif getattr(mod, 'op_kind', None) == 'up':
                Hs, Ws = Hs * mod.rate, Ws * mod.rate
        return lay, off

    def _build_forward_plan(self, B, H, W, mode, want_elem):
        """mode: 'eval' | 'train' | 'compress'"""
        m = self.model
        P = Plan(self, B, H, W, mode, want_elem)
        nH, nW = H // m.max_stride, W // m.max_stride
        P.im = P.f32(B, 3, H, W)
        P.lmb = P.f32(B)
        lay, kl_cols = self._latent_layout(B, nH, nW)
        P.layout = lay
        P.kl_partial = torch.zeros(B, kl_cols, dtype=torch.float32, device=self.device)
        P.z, P.kl_elem, P.sym, P.idx, P.noise = [], [], [], [], []
        if mode == 'compress':      # all layers' symbols / indexes in one buffer each: one D2H copy, one coder call
            total = sum(B * hw * zd for (hw, zd, _, _, _, _) in lay)
            P.sym_all, P.idx_all, P.sym_used = P.i32(total), P.i32(total), 0
            P.sym_host = torch.empty(total, dtype=torch.int32, pin_memory=True)
            P.idx_host = torch.empty(total, dtype=torch.int32, pin_memory=True)
            P.enc_layout = None
        self._embedding(P, P.lmb)
        feats = self._encoder(P, P.im)

        def latent_fn(P, blk, li, x, prior, geom):
            hw, zd, np_, off, Hs, Ws = lay[li]
            z = P.f32(B * hw, zd)
            kle = P.f32(B * hw, zd) if want_elem else None
            klp = P.kl_partial[:, off:]
            fuse = None
            if self.family == 'qarv' and mode in ('eval', 'compress') and self.npl and self.fuse_latent:
                # eval / compress: quantise + likelihood (+ symbols, indexes) as the epilogue of the posterior convolution
                sym = idx = None
                if mode == 'compress':
                    n_el = B * zd * Hs * Ws
                    sym = P.sym_all[P.sym_used:P.sym_used + n_el].view(B, zd, Hs, Ws)
                    idx = P.idx_all[P.sym_used:P.sym_used + n_el].view(B, zd, Hs, Ws)
                tab = self.w[id(blk)]['table']
                if tab is None and mode == 'compress':
                    raise ValueError('Uninitialized CDFs. Run update() first')      # CompressAI's message
                lat = N.LatentEpilogue()
                lat.prior, lat.scale_table, lat.n_scales = _ptr(prior), _ptr(tab), 0 if tab is None else tab.numel()
                lat.cdf_kind = N.CDF_ERFC if blk.discrete_gaussian.cdf_kind == 'erfc' else N.CDF_NORMAL
                lat.kl_partial, lat.kl_stride, lat.kl_elem = klp.data_ptr(), kl_cols, _ptr(kle)
                lat.sym, lat.idx = _ptr(sym), _ptr(idx)
                fuse = dict(lat=lat, z=z, keep=(prior, tab, kle, sym, idx))
            qm = self._posterior(P, blk, x, feats[blk.enc_key] if self.family == 'qarv' else feats[Hs], geom, fuse=fuse)
            P.z.append(z)
            P.kl_elem.append(kle)
            if fuse is not None and fuse.get('done'):
                P.keepalive = getattr(P, 'keepalive', []) + [fuse]
                if mode == 'compress':
                    P.sym_used += B * zd * Hs * Ws
                    P.sym.append(fuse['keep'][3])
                    P.idx.append(fuse['keep'][4])
                return z
            if self.family == 'rd':
                # continuous posterior: z = qm + qv * eps with eps ~ N(0,1) also in eval (rd/model.py:206-213)
                noise = P.f32(B * hw, zd)
                P.noise.append(noise)
                P.op('rd_latent', self.lib.lvae_rd_latent, _ptr(qm), _ptr(prior), _ptr(noise), _ptr(z),
                     klp.data_ptr(), kl_cols, _ptr(kle), B, hw, zd, keep=(qm, prior, z, kle),
                     meta=dict(kind='latent', elems=B * hw * zd, bytes=B * hw * zd * (28 - (0 if want_elem else 4))))
            elif mode == 'train':
                noise = P.f32(B * hw, zd)
                P.noise.append(noise)
                P.op('latent_train', self.lib.lvae_latent_train, _ptr(qm), _ptr(prior), _ptr(noise), _ptr(z),
                     klp.data_ptr(), kl_cols, _ptr(kle), B, hw, zd, keep=(qm, prior, z, kle),
                     meta=dict(kind='latent', elems=B * hw * zd, bytes=B * hw * zd * (20 + (4 if want_elem else 0))))
            else:
                sym = idx = None
                if mode == 'compress':
                    n_el = B * zd * Hs * Ws
                    sym = P.sym_all[P.sym_used:P.sym_used + n_el].view(B, zd, Hs, Ws)
                    idx = P.idx_all[P.sym_used:P.sym_used + n_el].view(B, zd, Hs, Ws)
                    P.sym_used += n_el
                    P.sym.append(sym)
                    P.idx.append(idx)
                tab = self.w[id(blk)]['table']
                if tab is None and mode == 'compress':
                    raise ValueError('Uninitialized CDFs. Run update() first')      # CompressAI's message
                cdf_kind = N.CDF_ERFC if blk.discrete_gaussian.cdf_kind == 'erfc' else N.CDF_NORMAL
                P.op('latent_eval', self.lib.lvae_latent_eval, _ptr(qm), _ptr(prior), _ptr(tab), 0 if tab is None else tab.numel(),
                     _ptr(z), klp.data_ptr(), kl_cols, _ptr(kle), _ptr(sym), _ptr(idx), B, hw, zd, cdf_kind, keep=(qm, prior, z, kle),
                     meta=dict(kind='latent', elems=B * hw * zd,
                               bytes=B * hw * zd * (16 + (4 if want_elem else 0) + (8 if mode == 'compress' else 0))))
            return z

        x_hat = self._top_down(P, feats, nH, nW, latent_fn, stop_at_flag=(mode == 'compress'))
        lossless = getattr(m, 'lossless', False)
        chw = 3 * H * W
        if lossless:                # GaussianNLLOutputNet: x_hat is the H/4 feature -> mean / log-scale images (NCHW)
            x_hat, P.out_ls = self._out_heads(P, x_hat, B, H, W)
        if mode != 'compress':
            P.x_hat = x_hat
            npi = self.lib.lvae_image_num_partials(chw)
            P.im_hat = P.f32(B, 3, H, W)
            pt, pi = P.f32(B, npi), P.f32(B, npi)
            P.stats = P.f32(4 + 3 * B)
            P.op('distortion', self.lib.lvae_image_distortion, _ptr(x_hat), _ptr(P.im), _ptr(P.im_hat), _ptr(pt), _ptr(pi),
                 B, chw, keep=(pt, pi))
            if lossless:            # the out-net loss is the NLL, not lambda * MSE: its per-image sums replace `pt` (lmb = 1)
                pt = P.f32(B, npi)
                P.op('nll', self.lib.lvae_nll_output, _ptr(x_hat), _ptr(P.out_ls), _ptr(P.im), _ptr(pt), B, chw, keep=(pt,))
            P.op('finalize', self.lib.lvae_rd_finalize, _ptr(P.kl_partial), kl_cols, kl_cols, _ptr(pt), _ptr(pi), npi,
                 _ptr(P.lmb), B, chw, _ptr(P.stats))
            P.stats_host = torch.empty(4 + 3 * B, dtype=torch.float32, pin_memory=True)
        elif lossless:              # the image's own residual stream (qresvae/model.py:81-86)
            tab = self.w['out_table']
            if tab is None:
                raise ValueError('Uninitialized CDFs. Run update() first')
            P.on_pm, P.on_idx, P.on_sym = P.f32(B * chw), P.i32(B * chw), P.i32(B * chw)
            P.on_idx_host = torch.empty(B * chw, dtype=torch.int32, pin_memory=True)
            P.on_sym_host = torch.empty(B * chw, dtype=torch.int32, pin_memory=True)
            P.op('outnet_codec', self.lib.lvae_outnet_codec, _ptr(x_hat), _ptr(P.out_ls), _ptr(P.im), _ptr(tab), tab.numel(),
                 _ptr(P.on_pm), _ptr(P.on_idx), _ptr(P.on_sym), B * chw, keep=(x_hat, tab))
        return P

    def _out_heads(self, P, feat, B, H, W):
        """GaussianNLLOutputNet's conv_mean / conv_scale (patch_upsample(C, 3, rate)) on the decoder's last feature
        [B*h*w, C] -> (p_mean, p_logscale), each NCHW [B, 3, H, W]"""
        on = self.model.out_net
        r = on.conv_mean.rate
        Hs, Ws, Cc = H // r, W // r, on.conv_mean[0].in_channels
        outs = []
        for head in (on.conv_mean, on.conv_scale):
            out = P.f32(B, 3, H, W)
            self._gemm(P, 'up', feat, (B, Hs, Ws, Cc, 1, 1, 0), self.w[id(head)], out, epi=N.EPI_SHUFFLE_NCHW, r=r)
            outs.append(out)
        return outs[0], outs[1]

    def _get_plan(self, key, builder):
        """Launch plans are cached per (batch, height, width, mode): each owns its activation buffers, pinned host mirrors
        and CUDA graphs, i.e. hundreds of MB at Kodak size.  The cache is an LRU of `max_plans` entries (LVAE_MAX_PLANS,
        default 6) so that a data set with many resolutions (CLIC, Tecnick through compress_file / self_evaluate) runs in
        bounded memory, like the reference; an evicted plan's graphs and buffers are freed before the new one is built."""
        P = self._plans.pop(key, None)
        if P is None:
            while len(self._plans) >= max(1, self.max_plans):
                old = self._plans.pop(next(iter(self._plans)))
                old.graphs = None
                del old
            with torch.cuda.device(self.device):
                P = builder()
        self._plans[key] = P            # most recently used last
        return P

    def _launch(self, P, seg=0):
        """Replay one plan segment on the current stream (through a CUDA graph once warmed up)."""
        N.launch_count += len(P.segments[seg])
        if not self.use_graphs:
            P.run_segment(seg, self._stream())
            return
        if P.graphs is None:
            P.graphs = [None] * len(P.segments)
            P.warm = [0] * len(P.segments)
        g = P.graphs[seg]
        if g is None:
            if P.warm[seg] < 1:           # first call runs eagerly (lazy module loading, error surfacing)
                P.warm[seg] += 1
                P.run_segment(seg, self._stream())
                return
            g = torch.cuda.CUDAGraph()
            cur = torch.cuda.current_stream(self.device)
            cur.synchronize()
            with torch.cuda.graph(g, capture_error_mode='thread_local'):
                P.run_segment(seg, self._stream())
            P.graphs[seg] = g
        g.replay()

    # ------------------------------------------------------------------ measurement hooks (bench.py)
    def forward_plan(self, B, H, W, mode='eval', want_elem=False):
        """The resident launch plan for one batch shape; fill P.im / P.lmb, then replay(P)."""
        self.refresh_weights()
        return self._get_plan((B, H, W, mode, want_elem), lambda: self._build_forward_plan(B, H, W, mode, want_elem))

    def replay(self, P, seg=0):
        with torch.cuda.device(self.device):
            self._launch(P, seg)

    def profile_ops(self, P, reps=3, seg=0):
        """Per-launch device time (CUDA events on the launching stream, eager launches, mean of `reps`)
        for every op of a plan segment -> list of (name, meta, ms)."""
        ops = P.segments[seg]
        out = []
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream(self.device)
            ev = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in ops]
                  for _ in range(reps)]
            for r in range(reps):
                for o, (e0, e1) in zip(ops, ev[r]):
                    e0.record(st)
                    rc = o.fn(*o.args, st.cuda_stream)
                    e1.record(st)
                    if rc != 0:
                        N.check(rc, o.name)
            st.synchronize()
            for i, o in enumerate(ops):
                ms = sum(ev[r][i][0].elapsed_time(ev[r][i][1]) for r in range(reps)) / reps
                out.append((o.name, o.meta, ms))
        return out

    # ------------------------------------------------------------------ public entry points
    @torch.no_grad()
    def run(self, im, lmb, mode='eval', want_elem=False, want_im_hat=False, check_range=True, noise=None):
        """im: [B,3,H,W] fp32 in [0,1], on the host (pinned for an async copy) or on the device; lmb: [B].
        noise: optional per-layer [B,zdim,h,w] tensors replacing the generator draws (qarv train mode: U(-.5,.5);
        rd: N(0,1)) -- how the tests feed the oracle's values."""
        self.refresh_weights()
        B, _, H, W = im.shape
        with torch.cuda.device(self.device):
            P = self._get_plan((B, H, W, mode, want_elem), lambda: self._build_forward_plan(B, H, W, mode, want_elem))
            P.im.copy_(im, non_blocking=True)
            P.lmb.copy_(lmb.to(torch.float32), non_blocking=True)
            rng = torch.aminmax(P.im) if check_range else None        # read back with the results: one sync
            for li, nz in enumerate(P.noise):      # same generator order as the reference: one draw per layer
                if noise is not None:
                    hw, zd, _, _, Hs, Ws = P.layout[li]
                    nz.view(B, Hs, Ws, zd).copy_(noise[li].to(self.device).permute(0, 2, 3, 1))
                elif self.family == 'rd':
                    nz.normal_()
                else:
                    nz.uniform_(-0.5, 0.5)
            self._launch(P)
            if mode == 'compress':
                strings = self._encode_strings(P)
                self._assert_range(rng)
                res = dict(strings=strings)
                if getattr(self.model, 'lossless', False):
                    res['out_strings'] = self._encode_outnet(P)
                return res
            P.stats_host.copy_(P.stats, non_blocking=True)
            torch.cuda.current_stream(self.device).synchronize()
            self._assert_range(rng)
            res = dict(stats=P.stats.clone(), stats_host=P.stats_host.numpy().copy(), x_hat=P.x_hat)
            if self.family == 'qres':     # per-layer rate log of HierarchicalVAE.forward (qresvae/model.py:551-557)
                res['kl_layers'] = [P.kl_partial[:, off:off + np_].sum(dim=1) for (_, _, np_, off, _, _) in P.layout]
            if want_im_hat:
                res['im_hat'] = P.im_hat.clone()
            if want_elem:
                res['x_hat'] = P.x_hat.clone()
                res['kl_elem'] = [self._nchw(k, B, l) for k, l in zip(P.kl_elem, P.layout)]
                res['z'] = [self._nchw(z, B, l) for z, l in zip(P.z, P.layout)]
            return res

    def _stream_slots(self, P, depth):
        """Per-plan staging for run_stream: `depth` device copies of the input batch, pinned result slots and events."""
        S = getattr(P, 'stream_slots', None)
        if S is None or len(S['stage']) != depth:
            n_lay = len(P.layout)
            S = dict(stage=[torch.empty_like(P.im) for _ in range(depth)],
                     dev=[torch.zeros(2 + n_lay, dtype=torch.float32, device=self.device) for _ in range(depth)],
                     host=[torch.empty(P.stats.numel(), dtype=torch.float32, pin_memory=True) for _ in range(depth)],
                     host_x=[torch.empty(2 + n_lay, dtype=torch.float32, pin_memory=True) for _ in range(depth)],
                     ready=[torch.cuda.Event() for _ in range(depth)], free=[torch.cuda.Event() for _ in range(depth)],
                     done=[torch.cuda.Event() for _ in range(depth)], n=0)
            P.stream_slots = S
        return S

    @torch.no_grad()
    def run_stream(self, items, mode='eval', depth=2, check_range=True):
        """Pipelined form of run(): `items` yields (im [B,3,H,W] fp32 in [0,1], lmb [B]); one dict(stats_host[, kl_layers_host])
        per item comes back, in order, `depth - 1` items late.  The host->device copy of item i+1 runs on a copy stream
        into a staging slot while the launch plan of item i executes; the results of item i travel to a pinned slot
        behind its plan, so every item still pays its own H2D and D2H copy but neither sits on the critical path.
        Same kernels, same plan, same numbers as run() (tests/test_gpu_model.py::test_forward_stream_equals_forward)."""
        assert mode in ('eval', 'train') and depth >= 1
        self.refresh_weights()
        from collections import deque
        inflight, cur_key = deque(), None
        with torch.cuda.device(self.device):
            cur = torch.cuda.current_stream(self.device)
            if getattr(self, '_copy_stream', None) is None:
                self._copy_stream = torch.cuda.Stream(self.device)
            cs = self._copy_stream
            for im, lmb in items:
                B, _, H, W = im.shape
                key = (B, H, W, mode, False)
                while inflight and (key != cur_key or len(inflight) >= depth):     # a new shape drains the pipeline first:
                    yield self._collect(*inflight.popleft(), check_range)          # its plan may evict the one in flight
                cur_key = key
                P = self._get_plan(key, lambda: self._build_forward_plan(B, H, W, mode, False))
                S = self._stream_slots(P, depth)
                k = S['n'] % depth
                S['n'] += 1
                if im.device.type == 'cpu':                  # H2D on the copy stream, into this item's staging slot
                    cs.wait_event(S['free'][k])
                    with torch.cuda.stream(cs):
                        S['stage'][k].copy_(im, non_blocking=True)
                        S['ready'][k].record(cs)
                    cur.wait_event(S['ready'][k])
                    P.im.copy_(S['stage'][k], non_blocking=True)
                    S['free'][k].record(cur)
                else:
                    P.im.copy_(im, non_blocking=True)
                P.lmb.copy_(lmb.to(torch.float32), non_blocking=True)
                dv = S['dev'][k]
                if check_range:
                    lo, hi = torch.aminmax(P.im)
                    dv[0].copy_(lo); dv[1].copy_(hi)
                for nz in P.noise:
                    if self.family == 'rd':
                        nz.normal_()
                    else:
                        nz.uniform_(-0.5, 0.5)
                self._launch(P)
                S['host'][k].copy_(P.stats, non_blocking=True)
                if self.family == 'qres':
                    for li, (_, _, np_, off, _, _) in enumerate(P.layout):
                        dv[2 + li].copy_(P.kl_partial[:, off:off + np_].sum(dim=1).mean(0))
                S['host_x'][k].copy_(dv, non_blocking=True)
                S['done'][k].record(cur)
                inflight.append((P, k))
            while inflight:
                yield self._collect(*inflight.popleft(), check_range)

    def _collect(self, P, k, check_range):
        S = P.stream_slots
        S['done'][k].synchronize()
        x = S['host_x'][k].numpy().copy()
        if check_range:
            self._assert_range((x[0], x[1]))
        res = dict(stats_host=S['host'][k].numpy().copy())
        if self.family == 'qres':
            res['kl_layers_mean_host'] = x[2:]
        return res

    @staticmethod
    def _assert_range(rng):
        # reference: assert 0 <= im.min() <= im.max() <= 1 (lvae/models/qarv/model.py:220)
        if rng is not None:
            lo, hi = float(rng[0]), float(rng[1])
            assert 0 <= lo <= hi <= 1, f'image values must lie in [0, 1], got [{lo}, {hi}]'

    @staticmethod
    def _nchw(t, B, lay):
        hw, zd, _, _, Hs, Ws = lay
        return t.view(B, Hs, Ws, zd).permute(0, 3, 1, 2).contiguous()

    # ---- host entropy coding
    def _tables(self, blk):
        return blk.discrete_gaussian.host_tables()

    def _encode_strings(self, P):
        """per latent layer: list (over the batch) of rANS byte strings.  Symbols / indexes arrive in NCHW
        order, the order CompressAI flattens them in (qarv/model.py:106-108).  All layers' symbols travel to the
        host in one pinned copy and the (layer, image) streams are coded in parallel by the C coder
        (lvae_rans_encode_streams)."""
        blocks = [b for b in self.model.dec_blocks if getattr(b, 'is_latent_block', False)]
        P.sym_host.copy_(P.sym_all, non_blocking=True)
        P.idx_host.copy_(P.idx_all, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        t0 = time.perf_counter()
        cdf, clen, coff = self._tables(blocks[0])
        for blk in blocks[1:]:      # one table set for the whole model (true for every registered model)
            assert blk.discrete_gaussian._quantized_cdf.shape == blocks[0].discrete_gaussian._quantized_cdf.shape
        if P.enc_layout is None:
            # streams sorted by size (largest first) for the thread pool; begin[] needs contiguous ranges, so the
            # stream table holds explicit (start, count) pairs through a begin array per stream
            streams = []
            off = 0
            for li, (sym, _) in enumerate(zip(P.sym, P.idx)):
                per = sym[0].numel()
                for b in range(P.B):
                    streams.append((li, b, off + b * per, per))
                off += sym.numel()
            P.enc_layout = streams
            P.enc_cap = [int(self.lib.lvae_rans_bound(per)) for (_, _, _, per) in streams]
            P.enc_out = np.empty(sum(P.enc_cap), dtype=np.uint8)
        streams = P.enc_layout
        sym_np, idx_np = P.sym_host.numpy(), P.idx_host.numpy()
        n = len(streams)
        # streams are laid out back to back in layer-major order, so begin[i] / begin[i+1] are simply cumulative
        begin = np.array([st[2] for st in streams] + [streams[-1][2] + streams[-1][3]], dtype=np.int64)
        out_begin = np.concatenate([[0], np.cumsum(P.enc_cap)]).astype(np.int64)
        out_len = np.zeros(n, dtype=np.int64)
        N.check(self.lib.lvae_rans_encode_streams(sym_np.ctypes.data, idx_np.ctypes.data, begin.ctypes.data, n,
                                                  cdf.ctypes.data, cdf.shape[1], clen.ctypes.data, coff.ctypes.data,
                                                  cdf.shape[0], P.enc_out.ctypes.data, out_begin.ctypes.data,
                                                  out_len.ctypes.data, self.coder_threads), 'rans_encode_streams')
        out = [[None] * P.B for _ in P.sym]
        for i, (li, b, _, _) in enumerate(streams):
            out[li][b] = P.enc_out[out_begin[i]:out_begin[i] + out_len[i]].tobytes()
        self.host_coder_s += time.perf_counter() - t0
        return out

    def _encode_outnet(self, P):
        """The lossless model's last stream: per image the 3*H*W residual symbols of the image itself against the out-net's
        128-scale tables (one rANS stream per image, coded in parallel)."""
        B = P.B
        P.on_sym_host.copy_(P.on_sym, non_blocking=True)
        P.on_idx_host.copy_(P.on_idx, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        t0 = time.perf_counter()
        cdf, clen, coff = self.model.out_net.discrete_gaussian.host_tables()
        per = P.on_sym_host.numel() // B
        cap = int(self.lib.lvae_rans_bound(per))
        out = np.empty(B * cap, dtype=np.uint8)
        begin = np.arange(B + 1, dtype=np.int64) * per
        out_begin = np.arange(B + 1, dtype=np.int64) * cap
        out_len = np.zeros(B, dtype=np.int64)
        sym_np, idx_np = P.on_sym_host.numpy(), P.on_idx_host.numpy()
        N.check(self.lib.lvae_rans_encode_streams(sym_np.ctypes.data, idx_np.ctypes.data, begin.ctypes.data, B,
                                                  cdf.ctypes.data, cdf.shape[1], clen.ctypes.data, coff.ctypes.data,
                                                  cdf.shape[0], out.ctypes.data, out_begin.ctypes.data,
                                                  out_len.ctypes.data, self.coder_threads), 'rans_encode_streams')
        self.host_coder_s += time.perf_counter() - t0
        return [out[out_begin[b]:out_begin[b] + out_len[b]].tobytes() for b in range(B)]

    def _build_decode_plan(self, B, nH, nW, sampling=False):
        m = self.model
        H, W = nH * m.max_stride, nW * m.max_stride
        P = Plan(self, B, H, W, 'sample' if sampling else 'decompress', False)
        P.lmb = P.f32(B)
        lay, _ = self._latent_layout(B, nH, nW)
        P.layout = lay
        P.z, P.idx, P.sym, P.prior, P.idx_host, P.sym_host = [], [], [], [], [], []
        self._embedding(P, P.lmb)

        def latent_fn(P, blk, li, x, prior, geom):
            hw, zd, np_, off, Hs, Ws = lay[li]
            z = P.f32(B * hw, zd)
            P.z.append(z)
            P.prior.append(prior)
            if sampling:
                P.cut()          # host decides per layer: given latent (copied into z) or prior sample
                return z
            idx, sym = P.i32(B, zd, Hs, Ws), P.i32(B, zd, Hs, Ws)
            P.idx.append(idx)
            P.sym.append(sym)
            P.idx_host.append(torch.empty(B * zd * Hs * Ws, dtype=torch.int32, pin_memory=True))
            P.sym_host.append(torch.empty(B * zd * Hs * Ws, dtype=torch.int32, pin_memory=True))
            tab = self.w[id(blk)]['table']
            if tab is None:
                raise ValueError('Uninitialized CDFs. Run update() first')
            P.op('prior_index', self.lib.lvae_latent_prior_index, _ptr(prior), _ptr(tab), tab.numel(), _ptr(idx),
                 B, hw, zd, keep=(prior, idx))
            P.cut()              # host: D2H idx -> rANS decode -> H2D sym
            P.op('dequant', self.lib.lvae_latent_dequant, _ptr(sym), _ptr(prior), _ptr(z), B, hw, zd, keep=(sym, z))
            return z

        x_hat = self._top_down(P, None, nH, nW, latent_fn)
        P.x_hat = x_hat
        if getattr(m, 'lossless', False):
            chw = 3 * H * W
            P.x_hat, P.out_ls = self._out_heads(P, x_hat, B, H, W)
            if not sampling:
                tab = self.w['out_table']
                if tab is None:
                    raise ValueError('Uninitialized CDFs. Run update() first')
                P.on_pm, P.on_idx, P.on_sym = P.f32(B * chw), P.i32(B * chw), P.i32(B * chw)
                P.on_idx_host = torch.empty(B * chw, dtype=torch.int32, pin_memory=True)
                P.on_sym_host = torch.empty(B * chw, dtype=torch.int32, pin_memory=True)
                P.on_im_hat = P.f32(B, 3, H, W)
                P.op('outnet_codec', self.lib.lvae_outnet_codec, _ptr(P.x_hat), _ptr(P.out_ls), 0, _ptr(tab), tab.numel(),
                     _ptr(P.on_pm), _ptr(P.on_idx), 0, B * chw, keep=(tab,))
                P.cut()              # host: D2H indexes -> rANS decode of the image residual -> H2D symbols
                P.op('outnet_decode', self.lib.lvae_outnet_decode, _ptr(P.on_sym), _ptr(P.on_pm), _ptr(P.on_im_hat), B * chw)
        return P

    @torch.no_grad()
    def decompress(self, lmb, strings, bhw, out_strings=None):
        """strings[li]: the byte string of latent layer li, or a list of B of them (one per image).  The B streams of a
        layer are decoded on a thread pool (lvae_rans_decode_streams); layers stay sequential, because the prior of
        layer i + 1 needs z_i (qarv/model.py:546-554)."""
        self.refresh_weights()
        B, nH, nW = bhw
        blocks = [b for b in self.model.dec_blocks if getattr(b, 'is_latent_block', False)]
        if self.decode_pingpong and B >= 4 and B % 2 == 0 and not getattr(self.model, 'lossless', False):
            return self._decompress_pingpong(lmb, strings, bhw, blocks)
        with torch.cuda.device(self.device):
            P = self._get_plan(('dec', B, nH, nW), lambda: self._build_decode_plan(B, nH, nW))
            P.lmb.copy_(lmb.to(torch.float32), non_blocking=True)
            stream = torch.cuda.current_stream(self.device)
            for li, blk in enumerate(blocks):
                self._launch(P, li)
                P.idx_host[li].copy_(P.idx[li].view(-1), non_blocking=True)
                stream.synchronize()
                per_layer = strings[li] if isinstance(strings[li], (list, tuple)) else [strings[li]]
                self._decode_layer_host(P, li, blk, per_layer)
                P.sym[li].copy_(P.sym_host[li].view_as(P.sym[li]), non_blocking=True)
            self._launch(P, len(blocks))
            if getattr(self.model, 'lossless', False):
                # qresvae/model.py:88-94: indexes from the out-net's scale head, the image residual from its own stream
                assert out_strings is not None and len(out_strings) == B
                P.on_idx_host.copy_(P.on_idx, non_blocking=True)
                stream.synchronize()
                cdf, clen, coff = self.model.out_net.discrete_gaussian.host_tables()
                idx_np, sym_np = P.on_idx_host.numpy(), P.on_sym_host.numpy()
                per = idx_np.size // B
                t0 = time.perf_counter()
                data = np.frombuffer(b''.join(out_strings), dtype=np.uint8)
                in_begin = np.concatenate([[0], np.cumsum([len(s_) for s_ in out_strings])]).astype(np.int64)
                begin = np.arange(B + 1, dtype=np.int64) * per
                N.check(self.lib.lvae_rans_decode_streams(data.ctypes.data, in_begin.ctypes.data, idx_np.ctypes.data,
                                                          begin.ctypes.data, B, cdf.ctypes.data, cdf.shape[1],
                                                          clen.ctypes.data, coff.ctypes.data, cdf.shape[0],
                                                          sym_np.ctypes.data, self.coder_threads), 'rans_decode_streams')
                self.host_coder_s += time.perf_counter() - t0
                P.on_sym.copy_(P.on_sym_host, non_blocking=True)
                self._launch(P, len(blocks) + 1)
                return P.on_im_hat.clone()
            return P.x_hat.clone().clamp_(min=-1.0, max=1.0).mul_(0.5).add_(0.5)

    def _decode_layer_host(self, P, li, blk, per_layer):
        """rANS-decode latent layer li of plan P's images on the host thread pool: P.idx_host[li] (table indexes, already on
        the host) + the images' byte strings -> P.sym_host[li]."""
        cdf, clen, coff = self._tables(blk)
        idx_np = P.idx_host[li].numpy().reshape(-1)
        sym_np = P.sym_host[li].numpy().reshape(-1)
        nB = P.B
        assert len(per_layer) == nB
        t0 = time.perf_counter()
        per = idx_np.size // nB
        data = np.frombuffer(b''.join(per_layer), dtype=np.uint8)
        in_begin = np.concatenate([[0], np.cumsum([len(s_) for s_ in per_layer])]).astype(np.int64)
        begin = (np.arange(nB + 1, dtype=np.int64) * per)
        N.check(self.lib.lvae_rans_decode_streams(data.ctypes.data, in_begin.ctypes.data, idx_np.ctypes.data,
                                                  begin.ctypes.data, nB, cdf.ctypes.data, cdf.shape[1],
                                                  clen.ctypes.data, coff.ctypes.data, cdf.shape[0],
                                                  sym_np.ctypes.data, self.coder_threads), 'rans_decode_streams')
        self.host_coder_s += time.perf_counter() - t0

    def _decompress_pingpong(self, lmb, strings, bhw, blocks):
        """Batched decode with the coder off the GPU's critical path: the batch is cut into two halves with a
        decode plan each; while the host decodes layer i of one half, the GPU runs the other half's segment, and the other way
        round -- layers stay sequential per image (the prior of layer i + 1 needs z_i, qarv/model.py:546-554).  One stream,
        one event per half: the host waits for `its` half only.  Kernels are batch-invariant, so the images come out bit-equal
        to the one-plan decode and to per-image decompress() (tests/test_gpu_model.py::test_pingpong_decode_...)."""
        B, nH, nW = bhw
        hb = B // 2
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device)
            halves, evs = [], []
            for h in range(2):
                P = self._get_plan(('dec', hb, nH, nW, h), lambda: self._build_decode_plan(hb, nH, nW))
                P.lmb.copy_(lmb[h * hb:(h + 1) * hb].to(torch.float32), non_blocking=True)
                halves.append(P)
                evs.append(torch.cuda.Event())
            for h, P in enumerate(halves):
                self._launch(P, 0)
                P.idx_host[0].copy_(P.idx[0].view(-1), non_blocking=True)
                evs[h].record(stream)
            for li, blk in enumerate(blocks):
                per_layer = strings[li]
                assert isinstance(per_layer, (list, tuple)) and len(per_layer) == B
                for h, P in enumerate(halves):
                    evs[h].synchronize()                     # this half's indexes are on the host; the other half's segment runs on
                    self._decode_layer_host(P, li, blk, per_layer[h * hb:(h + 1) * hb])
                    P.sym[li].copy_(P.sym_host[li].view_as(P.sym[li]), non_blocking=True)
                    self._launch(P, li + 1)
                    if li + 1 < len(blocks):
                        P.idx_host[li + 1].copy_(P.idx[li + 1].view(-1), non_blocking=True)
                        evs[h].record(stream)
            x = torch.cat([P.x_hat for P in halves], dim=0)
            return x.clamp_(min=-1.0, max=1.0).mul_(0.5).add_(0.5)

    @torch.no_grad()
    def sample(self, lmb, latents, bhw, t):
        self.refresh_weights()
        B, nH, nW = bhw
        with torch.cuda.device(self.device):
            P = self._get_plan(('smp', B, nH, nW), lambda: self._build_decode_plan(B, nH, nW, sampling=True))
            P.lmb.copy_(lmb.to(torch.float32), non_blocking=True)
            for li in range(len(P.z)):
                self._launch(P, li)
                hw, zd, _, _, Hs, Ws = P.layout[li]
                if latents[li] is None:
                    rn = torch.randn(B * hw, zd, device=self.device)
                    if self.family == 'rd':
                        N.check(self.lib.lvae_rd_sample(_ptr(P.prior[li]), _ptr(rn), t, _ptr(P.z[li]),
                                                        B, hw, zd, self._stream()), 'rd_sample')
                    else:
                        un = torch.empty(B * hw, zd, device=self.device).uniform_(-0.5, 0.5)
                        N.check(self.lib.lvae_latent_sample(_ptr(P.prior[li]), _ptr(rn), _ptr(un), t, _ptr(P.z[li]),
                                                            B, hw, zd, self._stream()), 'latent_sample')
                    N.launch_count += 1
                else:
                    assert tuple(latents[li].shape) == (B, zd, Hs, Ws)
                    P.z[li].view(B, Hs, Ws, zd).copy_(latents[li].to(self.device).permute(0, 2, 3, 1))
            self._launch(P, len(P.z))
            if getattr(self.model, 'lossless', False):
                # GaussianNLLOutputNet.sample (qresvae/model.py:46-56): mean + scale * t * N(0,1); not a hot path
                x = P.x_hat + torch.exp(P.out_ls) * (1.0 if t is None else t) * torch.randn_like(P.x_hat)
                return x.clamp_(min=-1.0, max=1.0).mul_(0.5).add_(0.5)
            return P.x_hat.clone().clamp_(min=-1.0, max=1.0).mul_(0.5).add_(0.5)
