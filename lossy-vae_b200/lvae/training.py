"""Differentiable execution of the rate-distortion path (training step; BASELINE configs 2 and 3).

`model.forward()` routes here when the model is in training mode and autograd is recording, so that
`loss = model(batch)['loss']; loss.backward(); optimizer.step()` of train-var-rate.py / train-fix-rate.py works
unchanged (reference lvae/trainer.py:255-300, qarv/model.py:317-363, qresvae/model.py:517-569).

Forward: the SAME liblvae_b200 kernels as the inference plans (dwconv+LN+AdaLN, tcgen05 GEMMs / fused MLP, fused latent
kernel), launched op by op through `torch.autograd.Function`s that keep only each op's inputs (activation checkpointing
at op granularity: ConvNeXt block, convolution, VDBlock, latent layer).

Backward -- STATE OF THIS ROUND, per op:
  * latent layer: native (`lvae_latent_train_bwd`);
  * ConvNeXt block: native -- the fc1 pre-activation is recomputed with the forward kernels (dwconv+LN kernel, tcgen05
    GEMM); both data gradients are tcgen05 GEMMs on transposed packed weights, GELU' applied in the epilogue of the first;
    both weight gradients are tcgen05 GEMMs contracting over the pixels (`lvae_split_planes_t_ex` + split-K
    `lvae_gemm_wgrad`; gelu(h) and the bias gradients ride in the operand splits); the dwconv + LayerNorm + AdaLN /
    affine part runs on the kernels of csrc/dwln_bwd.cu.  (Layers with fewer than 1024 pixels: `torch.mm` weight
    gradients; a few [C]-sized elementwise torch ops per block.)
  * convolutions of the qarv path (patch down / up, 1x1 heads with K-concat / residual, the 3x3 posterior head): native --
    data gradient = tcgen05 GEMM on the transposed packed weight, weight gradient = tcgen05 split-K GEMM over the pixels
    (tap by tap for the 3x3), bias gradient in the operand split (`TrainPath.conv_backward`); torch only re-arranges data;
  * qres-specific ops (VDBlocks, the GELU-fused / channel-padded z_proj convs): ATen autograd on a recomputed sub-graph.
Everything ATen here is a LIBRARY call (cuBLAS / cuDNN), not this repo's product; it is what the native backward
kernels of the next round replace, one op at a time, each against the gradient tests in tests/test_gpu_train.py
(`TrainPath.native_bwd = False` runs every backward through ATen: the cross-check of the native pieces).
There is no CPU path: every Function launches CUDA kernels of liblvae_b200.so.
"""
import contextlib
import math
import os

import torch
import torch.nn.functional as F

from . import _native as N
from .engine import Plan, _ptr
from .models import common


class _EagerPlan(Plan):
    """A Plan whose ops run as they are emitted (on the current stream): lets the op Functions reuse the engine's
    launch helpers (`_block`, `_gemm`, `_vdblock`) -- the inference code path -- one op at a time."""

    def __init__(self, engine, B):
        super().__init__(engine, B, 0, 0, 'train', True)
        self.cur_xg = None

    def op(self, name, fn, *args, keep=None, meta=None):
        rc = fn(*args, self.eng._stream())
        if rc != 0:
            N.check(rc, name)
        N.launch_count += 1


def _grad_of(fn, inputs, gout):
    """ATen autograd over a recomputed sub-graph: d fn(*inputs) . gout for every tensor input that needs it."""
    with torch.enable_grad():
        ins = [None if t is None else t.detach().requires_grad_(True) for t in inputs]
        out = fn(*ins)
        live = [t for t in ins if t is not None]
        grads = torch.autograd.grad(out, live, gout, allow_unused=True)
    it = iter(grads)
    out = [None if t is None else next(it) for t in ins]
    return [None if g is None else g.contiguous() for g in out]      # cuDNN hands back channels-last strided filters


# ------------------------------------------------------------------------------------------ ATen restatements (backward only)
def _dwln_aten(x, ada, dw_w, dw_b, ln_w, ln_b, k):
    """x NHWC [B,H,W,C]; ada [B,2C] = (shift | scale) or None (affine LN).  common.py:145-152 / timm ConvNeXtBlock."""
    C_ = x.shape[-1]
    y = F.conv2d(x.permute(0, 3, 1, 2), dw_w, dw_b, padding=(k - 1) // 2, groups=C_).permute(0, 2, 3, 1)
    if ada is None:
        return F.layer_norm(y, (C_,), ln_w, ln_b, eps=1e-6)
    y = F.layer_norm(y, (C_,), eps=1e-6)
    shift, scale = ada[:, None, None, :C_], ada[:, None, None, C_:]
    return y * (1 + scale) + shift


def _conv_aten(x, x1, res, w, b, cfg):
    a = x if x1 is None else torch.cat([x, x1], dim=-1)
    if cfg.get('nchw_in'):
        a_nchw = a
    else:
        if cfg.get('a_act'):
            a = F.gelu(a)
        a_nchw = a.permute(0, 3, 1, 2)
    y = F.conv2d(a_nchw, w, b, stride=cfg['stride'], padding=cfg['pad'])
    if cfg.get('gelu'):
        y = F.gelu(y)
    if cfg.get('r'):
        y = F.pixel_shuffle(y, cfg['r'])
    if cfg.get('nchw_out'):
        return y
    y = y.permute(0, 2, 3, 1)
    return y if res is None else y + res


def _vd_aten(x, x1, *p):
    a = x if x1 is None else torch.cat([x, x1], dim=-1)
    h = a.permute(0, 3, 1, 2)
    pad = (p[2].shape[-1] - 1) // 2
    h = F.conv2d(F.gelu(h), p[0], p[1])
    h = F.conv2d(F.gelu(h), p[2], p[3], padding=pad)
    h = F.conv2d(F.gelu(h), p[4], p[5], padding=pad)
    return F.conv2d(F.gelu(h), p[6], p[7]).permute(0, 2, 3, 1)


# ------------------------------------------------------------------------------------------ op Functions
class _BlockFn(torch.autograd.Function):
    """ConvNeXt block (AdaLN | affine LN).  x [B,H,W,C]; ada: THIS block's [B, 2C] column slice -- a view, with the row stride
    ada_total, of the [B, ada_total] matrix of all blocks' (shift | scale)
    rows (AdaLN) or None; params = conv_dw.weight, conv_dw.bias, fc1.weight, fc1.bias, fc2.weight, fc2.bias, gamma
    [, norm.weight, norm.bias]."""

    @staticmethod
    def forward(ctx, T, blk, x, ada, *params):
        B, H, W, C_ = x.shape
        out = torch.empty_like(x)
        T.P.ada = T.ada_full                       # the kernels address the full matrix (pointer, row stride, column offset)
        T.eng._block(T.P, blk, x, B, H, W, out=out)
        ctx.T, ctx.blk = T, blk
        ctx.save_for_backward(x, ada, *params)
        return out

    @staticmethod
    def backward(ctx, gout):
        T, blk = ctx.T, ctx.blk
        x, ada, *params = ctx.saved_tensors
        gout = gout.contiguous()
        if T.native_bwd and T.eng.npl:
            return (None, None) + T.block_backward(blk, x, ada, params, gout)
        C_, k = blk.dim, blk.kernel_size
        ada_s = ada

        def fn(x_, ada_, dw_w, dw_b, w1, b1, w2, b2, gamma, ln_w=None, ln_b=None):
            y = _dwln_aten(x_, ada_, dw_w, dw_b, ln_w, ln_b, k)
            y = F.linear(F.gelu(F.linear(y, w1, b1)), w2, b2)
            return x_ + y * gamma.reshape(-1)
        g = _grad_of(fn, [x, ada_s] + list(params), gout)
        return (None, None, g[0], g[1]) + tuple(g[2:])


class _ConvFn(torch.autograd.Function):
    """One convolution through lvae_gemm.  cfg: ks, stride, pad, epi, r (pixel shuffle), a_act, gelu, nchw_in (image
    patch embedding), nchw_out (the last pixel shuffle writes NCHW), pad_c (zero channels appended to x: z_proj of qres)."""

    @staticmethod
    def forward(ctx, T, went, cfg, x, x1, res, w, b):
        eng, P = T.eng, T.P
        if cfg.get('nchw_in'):                     # image -> space-to-depth operand (qarv/model.py:213-222 + patch conv)
            B, _, H, W = x.shape
            r = cfg['stride']
            Ho, Wo, m = H // r, W // r, T.model
            a = torch.empty(B * Ho * Wo, 3 * r * r, device=x.device)
            P.op('im2patch', eng.lib.lvae_image_to_patches, _ptr(x), _ptr(a), B, H, W, r, float(m.im_shift), float(m.im_scale))
            geom = (1, 1, B * Ho * Wo, 3 * r * r, 1, 1, 0)
        else:
            B, H, W, C0 = x.shape
            a = x
            if cfg.get('pad_c'):
                a = torch.empty(B, H, W, C0 + cfg['pad_c'], device=x.device)
                P.op('pad_z', eng.lib.lvae_pad_channels, _ptr(x), _ptr(a), B * H * W, C0, C0 + cfg['pad_c'])
                C0 += cfg['pad_c']
            ks, st, pad = cfg['ks'], cfg['stride'], cfg['pad']
            Ho, Wo = (H + 2 * pad - ks) // st + 1, (W + 2 * pad - ks) // st + 1
            geom = (B, H, W, C0, ks, st, pad)
        Nn, r = went['N'], cfg.get('r', 0)
        if cfg.get('nchw_out'):
            out = torch.empty(B, Nn // (r * r), Ho * r, Wo * r, device=x.device)
        elif r:
            out = torch.empty(B, Ho * r, Wo * r, Nn // (r * r), device=x.device)
        else:
            out = torch.empty(B, Ho, Wo, Nn, device=x.device)
        eng._gemm(P, cfg.get('name', 'conv'), a, geom, went, out, epi=cfg['epi'], a1=x1, C1=0 if x1 is None else x1.shape[-1],
                  res=res, r=r, a_act=cfg.get('a_act', 0))
        ctx.T, ctx.cfg, ctx.went = T, cfg, went
        ctx.save_for_backward(x, x1, res, w, b)
        return out

    @staticmethod
    def backward(ctx, gout):
        x, x1, res, w, b = ctx.saved_tensors
        cfg, m = ctx.cfg, ctx.T.model
        T = ctx.T
        if T.native_bwd and T.eng.npl and not (cfg.get('gelu') or cfg.get('a_act') or cfg.get('pad_c')):
            g = T.conv_backward(cfg, ctx.went, x, x1, res, w, b, gout.contiguous())
            return (None, None, None) + tuple(g)
        if cfg.get('nchw_in'):
            xin = x.add(m.im_shift).mul_(m.im_scale)
            g = _grad_of(lambda w_, b_: _conv_aten(xin, None, None, w_, b_, cfg), [w, b], gout)
            return (None, None, None, None, None, None, g[0], g[1])
        g = _grad_of(lambda *a: _conv_aten(*a, cfg), [x, x1, res, w, b], gout)
        return (None, None, None) + tuple(g)


class _VDFn(torch.autograd.Function):
    """qres VDBlock head without residual (qresvae/model.py:143-149); params = c1..c4 (weight, bias)."""

    @staticmethod
    def forward(ctx, T, wv, x, x1, *params):
        B, H, W, C0 = x.shape
        out = torch.empty(B, H, W, wv['c4']['N'], device=x.device)
        T.eng._vdblock(T.P, 'vd', wv, x, (B, H, W, C0), out, a1=x1, C1=0 if x1 is None else x1.shape[-1])
        ctx.T, ctx.wv = T, wv
        ctx.save_for_backward(x, x1, *params)
        return out

    @staticmethod
    def backward(ctx, gout):
        T = ctx.T
        if T.native_bwd and T.native_vd and T.eng.npl:
            x, x1, *params = ctx.saved_tensors
            return (None, None) + tuple(T.vd_backward(ctx.wv, x, x1, params, gout.contiguous()))
        g = _grad_of(_vd_aten, list(ctx.saved_tensors), gout)
        return (None, None) + tuple(g)


class _LatentFn(torch.autograd.Function):
    """z = qm + noise, kl = -gaussian_log_prob_mass(pm, pv, z) (entropy_coding.py:17-49); native both ways."""

    @staticmethod
    def forward(ctx, T, qm, prior, noise):
        B, H, W, zd = qm.shape
        z, kl = torch.empty_like(qm), torch.empty_like(qm)
        scratch = T.P.named('kl_partial', B * T.eng.lib.lvae_latent_num_partials(H * W, zd))
        T.P.op('latent_train', T.eng.lib.lvae_latent_train, _ptr(qm), _ptr(prior), _ptr(noise), _ptr(z), _ptr(scratch),
               scratch.numel() // B, _ptr(kl), B, H * W, zd)
        ctx.T = T
        ctx.save_for_backward(qm, prior, noise)
        return z, kl

    @staticmethod
    def backward(ctx, gz, gkl):
        qm, prior, noise = ctx.saved_tensors
        B, H, W, zd = qm.shape
        gz = None if gz is None else gz.contiguous()
        gkl = torch.zeros_like(qm) if gkl is None else gkl.contiguous()
        dqm, dprior = torch.empty_like(qm), torch.empty_like(prior)
        ctx.T.P.op('latent_train_bwd', ctx.T.eng.lib.lvae_latent_train_bwd, _ptr(qm), _ptr(prior), _ptr(noise), _ptr(gz),
                   _ptr(gkl), 0.0, _ptr(dqm), _ptr(dprior), B, H * W, zd)
        return None, dqm, dprior, None


# ------------------------------------------------------------------------------------------ the executor
class TrainPath:
    def __init__(self, model):
        self.model, self.eng = model, model.engine
        self.family = self.eng.family
        self.P = None
        # LVAE_TRAIN_NATIVE_BWD=0: every backward through ATen on recomputed sub-graphs (the cross-check of the native pieces)
        self.native_bwd = os.environ.get('LVAE_TRAIN_NATIVE_BWD', '1') != '0'
        self.native_wgrad = os.environ.get('LVAE_TRAIN_NATIVE_WGRAD', '1') != '0'
        # qres VDBlock heads: the native backward (vd_backward) is gradient-exact (tests/test_gpu_train.py) but SLOWER than
        # cuDNN's at the configs[2] shape (229 against 179 ms per step: the tap-by-tap 3x3 weight gradients are 9 transposing
        # splits + 9 split-K GEMMs per convolution), so it is opt-in (LVAE_TRAIN_NATIVE_VD=1) until the shifted-operand weight
        # gradient (one transposed plane set, 9 column offsets) replaces it
        self.native_vd = os.environ.get('LVAE_TRAIN_NATIVE_VD', '0') == '1'
        self.force_refresh = False      # GraphedTrainStep: the captured step must always re-pack the weights
        self.autograph_enabled = os.environ.get('LVAE_TRAIN_AUTOGRAPH', '1') != '0'
        self.autograph = AutoGraphedTrain(self)
        # weight gradients (operand splits + split-K GEMMs + their [C]-sized follow-ups) run on a second stream, forked from
        # and joined back into the stream of the backward inside every op's backward(): they depend only on the op's dY and
        # saved input, not on each other, and the layers behind the first two stages launch grids far smaller than the GPU
        # (LVAE_TRAIN_SIDE_STREAM=0: everything on one stream).  Fork / join are events, so a CUDA-graph capture of the step
        # records them as parallel branches.
        self.side_enabled = os.environ.get('LVAE_TRAIN_SIDE_STREAM', '1') != '0'
        self.side_stream = None         # created on first use
        # ConvNeXt block backward: the fc2 weight gradient, the fc1 weight gradient (+ its fp32 dwln recompute) and the depthwise filter
        # gradient each on a stream of their own (1 lane: 356, 2: 377, 3: 381-385 images/s at qarv 16 x 256^2)
        self.side_lanes = int(os.environ.get('LVAE_TRAIN_SIDE_LANES', '3'))
        # fc1 weight gradient: the A operand from the recomputed forward planes (lvae_planes_transpose) instead of an fp32 dwln recompute
        self.a_from_planes = os.environ.get('LVAE_TRAIN_A_FROM_PLANES', '1') != '0'
        self._side_used = False
        self._side_keep = []

    # ---- helpers
    @contextlib.contextmanager
    def _side(self, *keep, lane=0):
        """Run the enclosed launches on the side stream, ordered after everything the current stream has been given so far.
        `keep`: tensors of the current stream that the side work reads -- held until _join() so that the caching allocator
        cannot hand their memory to a later allocation of the main stream while the side stream still reads it."""
        if not self.side_enabled:
            yield
            return
        if self.side_stream is None:
            self.side_stream = torch.cuda.Stream(self.eng.device)
            self.side_streams = [self.side_stream] + [torch.cuda.Stream(self.eng.device) for _ in range(3)]
        st = self.side_streams[lane % max(1, min(self.side_lanes, len(self.side_streams)))]
        main = torch.cuda.current_stream(self.eng.device)
        ev = torch.cuda.Event()
        ev.record(main)
        st.wait_event(ev)
        self._side_used = (self._side_used or set()) | {st}
        self._side_keep.extend(keep)
        with torch.cuda.stream(st):
            yield

    def _join(self):
        """The current stream waits for the side stream (end of every backward())."""
        if self._side_used:
            for st in self._side_used:               # only the streams this backward() forked onto
                ev = torch.cuda.Event()
                ev.record(st)
                torch.cuda.current_stream(self.eng.device).wait_event(ev)
            self._side_used = False
        self._side_keep.clear()

    def split_ada(self, ada):
        """[B, ada_total] -> {id(block): its [B, 2C] column slice}.  torch.split, so that autograd concatenates the blocks'
        modulation gradients ONCE (SplitWithSizesBackward) -- slicing per block made every block's backward allocate a zero
        [B, ada_total] matrix, fill one slice, and autograd add 90 of them up."""
        self.ada_full = ada
        if ada is None:
            return {}
        blocks = sorted((b for b in self.eng.blocks if id(b) in self.eng.ada_off), key=lambda b: self.eng.ada_off[id(b)])
        sizes = [2 * b.dim for b in blocks]
        assert sum(sizes) == ada.shape[1] and all(self.eng.ada_off[id(b)] == sum(sizes[:i]) for i, b in enumerate(blocks))
        return {id(b): part for b, part in zip(blocks, ada.split(sizes, dim=1))}

    def _block(self, blk, x, ada):
        ada = ada.get(id(blk)) if isinstance(ada, dict) else ada
        ps = [blk.conv_dw.weight, blk.conv_dw.bias, blk.mlp.fc1.weight, blk.mlp.fc1.bias, blk.mlp.fc2.weight,
              blk.mlp.fc2.bias, blk.gamma]
        if isinstance(blk, common.ConvNeXtBlockLN):
            ps += [blk.norm.weight, blk.norm.bias]
        return _BlockFn.apply(self, blk, x, ada, *ps)

    def _conv(self, conv, went, x, x1=None, res=None, **cfg):
        cfg.setdefault('ks', conv.kernel_size[0])
        cfg.setdefault('stride', conv.stride[0])
        cfg.setdefault('pad', conv.padding[0])
        cfg.setdefault('epi', N.EPI_BIAS_RES if res is not None else N.EPI_BIAS)
        return _ConvFn.apply(self, went, cfg, x, x1, res, conv.weight, conv.bias)

    def _vd(self, vd, wv, x, x1=None):
        ps = []
        for k in ('c1', 'c2', 'c3', 'c4'):
            ps += [getattr(vd, k).weight, getattr(vd, k).bias]
        return _VDFn.apply(self, wv, x, x1, *ps)

    # Data-gradient GEMMs run in the 2-plane bf16 mode whatever the forward mode is: their A operand is a GRADIENT, whose
    # magnitude (1e-3 ... 1e-9 at the default init, where the layer scale is 1e-6) is far outside what fp16 planes hold;
    # bf16 planes have fp32's exponent range, and 16 significand bits are ample for a gradient (cuDNN's default for the
    # reference's convolutions is TF32: 11 bits).
    DGRAD_PREC = N.PREC_BF16X3

    def _transposed(self, w2d, row_scale=None):
        """packed operand planes of a transposed (and possibly row-scaled) weight for the data-gradient GEMMs.  With at
        least 64 rows (never the exact-fp32 small-K path, which reads the fp32 matrix) the planes of w2d^T come straight from
        the transposing split (lvae_split_planes_t: same bf16 hi / lo roundings as lvae_split_planes) -- one launch instead of a
        transposed copy + a split; row_scale [N]: rows of w2d scaled first (the layer scale folded into fc2's data gradient)."""
        if row_scale is not None:
            w2d = row_scale[:, None] * w2d
        Nr, Kc = w2d.shape
        if Nr >= 64 and Nr % 2 == 0 and self.DGRAD_PREC == N.PREC_BF16X3:
            w2d = w2d.contiguous()
            pl = [torch.empty(Kc * Nr, dtype=torch.bfloat16, device=w2d.device) for _ in range(2)]
            self.P.op('split_t', self.eng.lib.lvae_split_planes_t_ex, _ptr(w2d), _ptr(pl[0]), _ptr(pl[1]), Nr, Kc, 0, 0)
            return dict(w=w2d, bias=None, N=Kc, K=Nr, planes=pl)       # `w` only keeps the descriptor's pointer non-null
        return self.eng._pack_gemm_weight(w2d.t().contiguous(), None, prec=self.DGRAD_PREC)

    def _t_planes(self, name, x2d, act=0, colsum=None):
        """[P, C] fp32 -> two K-major bf16 planes [C, P] in a scratch buffer (the operand format of lvae_gemm_wgrad);
        act = 1: planes of gelu(x); colsum [C] (zeroed): += column sums of x (the bias gradient when x is a dY)"""
        P_, C_ = x2d.shape
        buf = self.P.named(name, 2 * P_ * C_, dtype=torch.bfloat16)
        p0, p1 = buf[:P_ * C_], buf[P_ * C_:2 * P_ * C_]
        self.P.op('split_t', self.eng.lib.lvae_split_planes_t_ex, _ptr(x2d), _ptr(p0), _ptr(p1), P_, C_, act, _ptr(colsum))
        return p0, p1

    def _wgrad(self, dy_t, x_t, n_out, k_in, P_):
        dw = torch.empty(n_out, k_in, device=self.eng.device)
        self.P.op('wgrad', self.eng.lib.lvae_gemm_wgrad, _ptr(dy_t[0]), _ptr(dy_t[1]), _ptr(x_t[0]), _ptr(x_t[1]), _ptr(dw),
                  n_out, k_in, P_)
        return dw

    def block_backward(self, blk, x, ada, params, gout):
        """Gradients of one ConvNeXt block w.r.t. (x, ada, *params).  The two GEMM data gradients and the recomputation
        of the fc1 pre-activation run on the tcgen05 GEMM; see the module docstring for what is still ATen."""
        eng, P = self.eng, self.P
        dw_w, dw_b, w1, b1, w2, b2, gamma = params[:7]
        ln = params[7:]
        B, H, W, C_ = x.shape
        M, hid, k = B * H * W, blk.hidden, blk.kernel_size
        wb = eng.w[id(blk)]
        off = 0                                     # `ada` is this block's slice: its data pointer already carries the offset
        gam = gamma.detach().reshape(-1)
        # recompute: a = AdaLN(LN(dwconv(x))) as operand planes, h = a W1^T + b1 (pre-activation)
        P.ada = ada
        A = [P.named(f'scratch_a{i}', M * C_, dtype=torch.bfloat16) for i in range(eng.npl)]
        ap = [_ptr(t) for t in A] + [0] * (3 - eng.npl)
        P.op('dwln', eng.lib.lvae_dwconv_ln_adaln_planes, _ptr(x), _ptr(wb['dw_w']), _ptr(wb['dw_b']), _ptr(ada),
             eng.ada_total, off, _ptr(wb.get('ln_w')), _ptr(wb.get('ln_b')), ap[0], ap[1], ap[2], eng.pfmt, B, H, W, C_, k)
        h = torch.empty(M, hid, device=x.device)
        eng._gemm(P, 'fc1.re', None, (1, 1, M, C_, 1, 1, 0), wb['fc1'], h, epi=N.EPI_BIAS, a_planes=A)
        go = gout.reshape(M, C_)
        # fc2: out = x + gamma * (g W2^T + b2),  g = gelu(h)
        tc_wgrad = self.native_wgrad and M % 8 == 0 and M >= 1024
        lib = eng.lib
        c = torch.empty(M, C_, device=x.device)
        c_ready = None
        with self._side(go, h, gam, c, x):
            if self.side_enabled:       # the conv output that the LayerNorm backward (last part, main stream) needs: first thing here
                P.op('dwconv', lib.lvae_dwconv, _ptr(x), _ptr(wb['dw_w']), _ptr(wb['dw_b']), 0, _ptr(c), B, H, W, C_, k, 0)
                c_ready = torch.cuda.Event()
                c_ready.record(self.side_stream)
            if tc_wgrad:                                            # tcgen05, split over the pixels (csrc/wgrad.cu)
                db2_raw = torch.zeros(C_, device=x.device)          # bias gradient rides in the operand split of dout
                go_t = self._t_planes('wg_a', go, colsum=db2_raw)
                dw2_raw = self._wgrad(go_t, self._t_planes('wg_b', h, act=1), C_, hid, M)
            else:                                                   # tiny layers: cuBLAS fp32  [C, hid]
                dw2_raw = go.t().mm(F.gelu(h))
                db2_raw = go.sum(0)
            d_w2 = gam[:, None] * dw2_raw
            d_b2 = gam * db2_raw
            d_gamma = ((w2.detach() * dw2_raw).sum(1) + b2.detach() * db2_raw).reshape(gamma.shape)
        # dh = ((dout * gamma) W2) * gelu'(h): GELU' applied in the GEMM's epilogue
        dh = torch.empty(M, hid, device=x.device)
        # dh leaves the epilogue twice: fp32 (for the transposing split of the fc1 weight gradient) and as the two bf16 operand
        # planes that the fc1 data-gradient GEMM reads -- no split pass over [M, hid] in between
        dh_pl = [P.named(f'dh_pl{i}', M * hid, dtype=torch.bfloat16)[:M * hid] for i in range(2)] if hid % 8 == 0 else None
        eng._gemm(P, 'fc2.dgrad', go, (1, 1, M, C_, 1, 1, 0), self._transposed(w2.detach(), row_scale=gam), dh,
                  epi=N.EPI_GELU_BWD, res=h, prec=self.DGRAD_PREC, out_planes=dh_pl)
        # fc1: h = a W1^T + b1;  a is needed for the weight gradient only: one more (fp32) dwln launch, on the side stream
        from_planes = tc_wgrad and eng.npl == 2 and self.a_from_planes
        a32 = None if from_planes else torch.empty(M, C_, device=x.device)
        with self._side(dh, a32, x, ada, A, lane=1):
            if from_planes:
                # the operand planes recomputed at the top (value = hi + lo, 22 significand bits) re-split into K-major bf16 planes:
                # no second dwln launch; the planes stay valid until the next block's backward (which starts after the join)
                buf = P.named('wg_c', 2 * M * C_, dtype=torch.bfloat16)
                a_t = (buf[:M * C_], buf[M * C_:2 * M * C_])
                P.op('planes_t', eng.lib.lvae_planes_transpose, _ptr(A[0]), _ptr(A[1]), eng.pfmt, _ptr(a_t[0]), _ptr(a_t[1]), M, C_)
            else:
                P.op('dwln', eng.lib.lvae_dwconv_ln_adaln, _ptr(x), _ptr(wb['dw_w']), _ptr(wb['dw_b']), _ptr(ada), eng.ada_total, off,
                     _ptr(wb.get('ln_w')), _ptr(wb.get('ln_b')), _ptr(a32), B, H, W, C_, k)
            if tc_wgrad:
                d_b1 = torch.zeros(hid, device=x.device)
                d_w1 = self._wgrad(self._t_planes('wg_d', dh, colsum=d_b1), a_t if from_planes else self._t_planes('wg_c', a32), hid, C_, M)
            else:
                d_w1 = dh.t().mm(a32)
                d_b1 = dh.sum(0)
        da = torch.empty(M, C_, device=x.device)
        eng._gemm(P, 'fc1.dgrad', None if dh_pl else dh, (1, 1, M, hid, 1, 1, 0), self._transposed(w1.detach()), da, epi=N.EPI_BIAS,
                  prec=self.DGRAD_PREC, a_planes=dh_pl)
        del dh, h, a32
        # dwconv + LayerNorm + modulation: recompute the conv output, LayerNorm / modulation backward, filter gradient,
        # data gradient (transposed conv + the residual branch's gradient) -- csrc/dwln_bwd.cu
        if c_ready is None:
            P.op('dwconv', lib.lvae_dwconv, _ptr(x), _ptr(wb['dw_w']), _ptr(wb['dw_b']), 0, _ptr(c), B, H, W, C_, k, 0)
        else:
            torch.cuda.current_stream(x.device).wait_event(c_ready)
        dmod, dc = torch.empty(1 if ln else B, 2 * C_, device=x.device), torch.empty(M, C_, device=x.device)
        P.op('ln_mod_bwd', lib.lvae_ln_mod_bwd, _ptr(c), _ptr(da), _ptr(ada), eng.ada_total, off, _ptr(wb.get('ln_w')),
             _ptr(dc), _ptr(dmod), B, H * W, C_)
        with self._side(dc, lane=2):
            dwp, d_dwb = torch.empty(k * k, C_, device=x.device), torch.empty(C_, device=x.device)
            P.op('dwconv_wgrad', lib.lvae_dwconv_wgrad, _ptr(dc), _ptr(x), _ptr(dwp), _ptr(d_dwb), B, H, W, C_, k)
            d_dww = dwp.t().reshape(C_, 1, k, k)
        dx = torch.empty_like(x)
        P.op('dwconv_dgrad', lib.lvae_dwconv, _ptr(dc), _ptr(wb['dw_w']), 0, _ptr(gout), _ptr(dx), B, H, W, C_, k, 1)
        self._join()
        if ln:
            return (dx, None, d_dww, d_dwb, d_w1, d_b1, d_w2, d_b2, d_gamma, dmod[0, C_:].clone(), dmod[0, :C_].clone())
        return (dx, dmod, d_dww, d_dwb, d_w1, d_b1, d_w2, d_b2, d_gamma)

    # ---- convolutions (patch down / up, 1x1 heads with K-concat and residual, the 3x3 posterior head): native backward
    def _mm_grad(self, name, g2d, a2d, colsum=None, act=0):
        """g2d^T @ act(a2d) -> [N, K] (weight gradient of a linear map; act = 1: GELU applied to a2d as it is split) + optional
        column sums of g2d (bias gradient): tcgen05 split-K GEMM over the pixels when there are enough of them, cuBLAS fp32
        for the tiny layers."""
        M, Nn = g2d.shape
        K = a2d.shape[1]
        if self.native_wgrad and M % 8 == 0 and M >= 1024:
            return self._wgrad(self._t_planes('wg_a', g2d.contiguous(), colsum=colsum), self._t_planes('wg_b', a2d.contiguous(), act=act), Nn, K, M)
        if colsum is not None:
            colsum.add_(g2d.sum(0))
        return g2d.t().mm(F.gelu(a2d) if act else a2d)

    def _conv3_grads(self, G4, X2, w4, act=0):
        """3x3 / stride 1 / pad 1 convolution y = conv(act(X), w4) + b with G4 = dL/dy [B, H, W, N], X2 [M, C] (its input
        BEFORE act), w4 [N, C, 3, 3] -> (d act(X) [M, C], dw4, db).  Data gradient: the same im2col GEMM of G with the
        flipped, transposed filter; weight gradient tap by tap: dW[:, :, ky, kx] = shift(G, ky - 1, kx - 1)^T @ act(X)."""
        eng = self.eng
        B, H, W_, Nn = G4.shape
        M, C0 = X2.shape
        dev = G4.device
        wd = w4.detach().flip(2, 3).permute(1, 2, 3, 0).reshape(C0, 9 * Nn).contiguous()      # [c, (ky, kx, n)], flipped
        dxm = torch.empty(M, C0, device=dev)
        eng._gemm(self.P, 'conv3.dgrad', G4, (B, H, W_, Nn, 3, 1, 1), eng._pack_gemm_weight(wd, None, prec=self.DGRAD_PREC),
                  dxm, epi=N.EPI_BIAS, prec=self.DGRAD_PREC)
        tc = self.native_wgrad and M % 8 == 0 and M >= 1024
        with self._side(G4, X2):
            x_t = self._t_planes('wg_b', X2.contiguous(), act=act) if tc else None
            Xa = None if tc else (F.gelu(X2) if act else X2)
            Gp = F.pad(G4, (0, 0, 1, 1, 1, 1))
            dw4 = torch.empty(Nn, C0, 3, 3, device=dev)
            db = G4.reshape(M, Nn).sum(0)
            ready = None
            if self.side_enabled:
                ready = torch.cuda.Event()
                ready.record(torch.cuda.current_stream(dev))
        # the nine taps are independent: round-robin over the side lanes, each lane with its own scratch planes for shift(G)
        lanes = max(1, self.side_lanes) if self.side_enabled else 1
        for tap in range(9):
            ky, kx = divmod(tap, 3)
            with self._side(Gp, x_t, Xa, dw4, lane=tap % lanes):
                if ready is not None:
                    torch.cuda.current_stream(dev).wait_event(ready)
                Gt = Gp[:, 2 - ky:2 - ky + H, 2 - kx:2 - kx + W_, :].reshape(M, Nn)
                if tc:
                    dw4[:, :, ky, kx] = self._wgrad(self._t_planes(f'wg_a{tap % lanes}', Gt.contiguous()), x_t, Nn, C0, M)
                else:
                    dw4[:, :, ky, kx] = Gt.t().mm(Xa)
        return dxm, dw4, db

    def vd_backward(self, wv, x, x1, params, gout):
        """Gradients (x, x1, c1.w, c1.b, ..., c4.w, c4.b) of a qres VDBlock head (qresvae/model.py:118-149, residual = False):
        out = c4(gelu(c3(gelu(c2(gelu(c1(gelu(cat(x, x1))))))))).  The pre-activations are recomputed in fp32 with the
        forward GEMM (GELU applied to the operand as it is read); every data gradient is a tcgen05 GEMM, every weight
        gradient a tcgen05 split-K GEMM over the pixels; GELU' is ATen's elementwise gelu_backward (no cuDNN anywhere)."""
        eng, P = self.eng, self.P
        w1, b1, w2, b2, w3, b3, w4, b4 = params
        B, H, W_, C0 = x.shape
        M = B * H * W_
        C1 = 0 if x1 is None else x1.shape[-1]
        hid, ks = wv['c1']['N'], wv['c2']['ks']
        dev = x.device
        pad = (ks - 1) // 2
        # ---- recompute the three pre-activations (fp32)
        h1 = torch.empty(B, H, W_, hid, device=dev)
        eng._gemm(P, 'vd.c1.re', x, (B, H, W_, C0, 1, 1, 0), wv['c1'], h1, a1=x1, C1=C1, a_act=1)
        h2 = torch.empty_like(h1)
        eng._gemm(P, 'vd.c2.re', h1, (B, H, W_, hid, ks, 1, pad), wv['c2'], h2, a_act=1)
        h3 = torch.empty_like(h1)
        eng._gemm(P, 'vd.c3.re', h2, (B, H, W_, hid, ks, 1, pad), wv['c3'], h3, a_act=1)
        gelu_bwd = torch.ops.aten.gelu_backward
        # ---- c4 (1x1): out = gelu(h3) W4^T + b4
        G = gout.reshape(M, -1)
        db4 = torch.zeros(G.shape[1], device=dev)
        with self._side(G, h3, db4):
            dw4 = self._mm_grad('vd.c4.wgrad', G, h3.view(M, hid), db4, act=1).reshape(w4.shape)
        dh3 = gelu_bwd(self._dgrad('vd.c4.dgrad', G, wv['c4']['w']), h3.view(M, hid))
        # ---- c3, c2 (3x3 | 1x1)
        def mid(dh, h_in, went, w):
            if ks == 3:
                dg, dw, db = self._conv3_grads(dh.view(B, H, W_, hid), h_in.view(M, hid), w, act=1)
            else:
                db = torch.zeros(hid, device=dev)
                with self._side(dh, h_in, db):
                    dw = self._mm_grad('vd.mid.wgrad', dh, h_in.view(M, hid), db, act=1).reshape(w.shape)
                dg = self._dgrad('vd.mid.dgrad', dh, went['w'])
            return gelu_bwd(dg, h_in.view(M, hid)), dw, db
        dh2, dw3, db3 = mid(dh3, h2, wv['c3'], w3)
        dh1, dw2, db2 = mid(dh2, h1, wv['c2'], w2)
        # ---- c1 (1x1 on gelu(cat(x, x1)))
        db1 = torch.zeros(hid, device=dev)
        x2 = x.reshape(M, C0)
        x12 = None if x1 is None else x1.reshape(M, C1)
        with self._side(dh1, x2, x12, db1):
            dwa = self._mm_grad('vd.c1.wgrad', dh1, x2, db1, act=1)
            if x1 is not None:
                dw1 = torch.cat([dwa, self._mm_grad('vd.c1.wgrad1', dh1, x12, act=1)], dim=1).reshape(w1.shape)
            else:
                dw1 = dwa.reshape(w1.shape)
        dga = self._dgrad('vd.c1.dgrad', dh1, wv['c1']['w'])          # [M, C0 + C1]
        if x1 is not None:
            dx = gelu_bwd(dga[:, :C0].contiguous(), x2).view(x.shape)
            dx1 = gelu_bwd(dga[:, C0:].contiguous(), x12).view(x1.shape)
        else:
            dx, dx1 = gelu_bwd(dga, x2).view(x.shape), None
        self._join()
        return dx, dx1, dw1, db1, dw2, db2, dw3, db3, dw4, db4

    def _dgrad(self, name, g2d, w2d):
        """g2d [M, N] @ w2d [N, K] -> [M, K]: the data gradient of a linear map, on the tcgen05 GEMM (2-plane bf16)."""
        M, Nn = g2d.shape
        if Nn % 8:                                          # qres z_dims 14 / 12 / 10: a contraction over < 16 channels
            return g2d.mm(w2d)
        out = torch.empty(M, w2d.shape[1], device=g2d.device)
        self.eng._gemm(self.P, name, g2d.contiguous(), (1, 1, M, Nn, 1, 1, 0), self._transposed(w2d), out, epi=N.EPI_BIAS, prec=self.DGRAD_PREC)
        return out

    def conv_backward(self, cfg, went, x, x1, res, w, b, gout):
        """Gradients (x, x1, res, w, b) of one _ConvFn op without ATen's convolution_backward: every convolution of the
        qarv path is a linear map on a (re-arranged) pixel matrix, so its data gradient is one tcgen05 GEMM on the
        transposed packed weight and its weight gradient one tcgen05 GEMM contracting over the pixels; what torch still
        does here is data movement (patch / pixel-shuffle re-arrangement views, copies, small pads)."""
        eng, m = self.eng, self.model
        ks, st, r = cfg['ks'], cfg['stride'], cfg.get('r', 0)
        Wp = went['w']                                     # packed [N, K] fp32 on the device, K order (ky, kx, c) | segment 1
        Nn, K = Wp.shape
        dev = gout.device
        # ---- G: [M, N] in the packed column order of the forward GEMM
        if r:
            Co = Nn // (r * r)
            if cfg.get('nchw_out'):
                B, _, Hr, Wr = gout.shape
                G = gout.view(B, Co, Hr // r, r, Wr // r, r).permute(0, 2, 4, 3, 5, 1)
            else:
                B, Hr, Wr, _ = gout.shape
                G = gout.view(B, Hr // r, r, Wr // r, r, Co).permute(0, 1, 3, 2, 4, 5)
            G = G.reshape(-1, Nn)
        else:
            G = gout.reshape(-1, Nn)
        M = G.shape[0]
        db_p = torch.zeros(Nn, device=dev) if b is not None else None
        dx = dx1 = None
        if ks == 3:
            # 3x3, stride 1, pad 1 (posterior head)
            B, H, W_, C0 = x.shape
            dxm, dw4, db = self._conv3_grads(gout, x.reshape(M, C0), w)
            self._join()
            return dxm.view(B, H, W_, C0), None, None, dw4, (db if b is not None else None)
        # ---- A: the forward operand matrix [M, K] (re-arranged input), and the way back for its gradient
        if cfg.get('nchw_in'):
            B, _, H, W_ = x.shape
            A = torch.empty(M, K, device=dev)
            self.P.op('im2patch', eng.lib.lvae_image_to_patches, _ptr(x), _ptr(A), B, H, W_, st, float(m.im_shift), float(m.im_scale))
            with self._side(G, A, db_p):
                dWp = self._mm_grad('down0.wgrad', G, A, db_p)
        else:
            B, H, W_, C0 = x.shape
            if ks > 1:                                     # non-overlapping patches: space-to-depth view
                Ho, Wo = H // ks, W_ // ks
                A = x.view(B, Ho, ks, Wo, ks, C0).permute(0, 1, 3, 2, 4, 5).reshape(M, K)
            else:
                A = x.reshape(M, C0)
            with self._side(G, A, db_p, x1):
                if x1 is not None:
                    C1 = x1.shape[-1]
                    dWp = torch.cat([self._mm_grad('conv.wgrad', G, A, db_p), self._mm_grad('conv.wgrad1', G, x1.reshape(M, C1))], dim=1)
                else:
                    dWp = self._mm_grad('conv.wgrad', G, A, db_p)
            dA = self._dgrad('conv.dgrad', G, Wp)          # [M, K]
            if x1 is not None:
                dx = dA[:, :C0].reshape(x.shape)
                dx1 = dA[:, C0:].reshape(x1.shape)
            else:
                if ks > 1:
                    dx = dA.view(B, Ho, Wo, ks, ks, C0).permute(0, 1, 3, 2, 4, 5).reshape(x.shape)
                else:
                    dx = dA.view(x.shape)
        # ---- packed [N, (ky, kx, c) | c1] -> the conv weight's layout
        with self._side():
            if r:                                          # pixel shuffle: packed row (i r + j) Co + c  <-  reference row c r r + i r + j
                Co = Nn // (r * r)
                perm = torch.arange(Nn, device=dev).reshape(Co, r * r).t().reshape(-1)
                dW_ref = torch.empty_like(dWp)
                dW_ref[perm] = dWp
                dWp = dW_ref
                if db_p is not None:
                    db_ref = torch.empty_like(db_p)
                    db_ref[perm] = db_p
                    db_p = db_ref
            Cin = w.shape[1]
            if ks > 1:
                dw4 = dWp.view(Nn, ks, ks, Cin).permute(0, 3, 1, 2).contiguous()
            else:
                dw4 = dWp.reshape(w.shape)
        self._join()
        return dx, dx1, (gout if res is not None else None), dw4, db_p

    # ---- lambda embedding (tiny; ATen both ways): qarv/model.py:280-287, common.py:101-107,150
    def _ada(self, lmb):
        m, eng = self.model, self.eng
        if not eng.ada_total:
            return None
        scaled = torch.log(lmb) * m._sin_period / math.log(m.MAX_LMB)
        freqs = eng.w['freqs']
        args = scaled.view(-1, 1) * freqs.view(1, -1)
        e = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
        e = F.linear(e, m.lmb_embedding[0].weight, m.lmb_embedding[0].bias)
        e = F.linear(F.gelu(e), m.lmb_embedding[2].weight, m.lmb_embedding[2].bias)
        blocks = [b for b in eng.blocks if id(b) in eng.ada_off]
        w_all = torch.cat([b.embedding_layer[1].weight for b in blocks], 0)
        b_all = torch.cat([b.embedding_layer[1].bias for b in blocks], 0)
        return F.linear(F.gelu(e), w_all, b_all).contiguous()

    # ---- the pass
    def forward(self, im, lmb, noise=None):
        """im [B,3,H,W] on the device in [0,1]; lmb [B]; noise: optional per-layer [B,zdim,h,w] U(-.5,.5).
        Returns dict(x_hat [B,3,H,W] (graph-attached), kl [per layer [B,h,w,zdim] graph-attached], z [per layer])."""
        m, eng = self.model, self.eng
        eng.refresh_weights(force=self.force_refresh)
        B, _, H, W = im.shape
        if self.P is None or self.P.B != B or self.P.dev != eng.device:
            self.P = _EagerPlan(eng, B)
        qres = self.family == 'qres'
        with torch.cuda.device(eng.device), torch.autocast('cuda', enabled=False):
            ada = self.split_ada(self._ada(lmb))
            # ---------------- bottom-up
            feats, x, s, first = {}, im, 1, True
            for mod in m.encoder.enc_blocks:
                kind = getattr(mod, 'op_kind', None)
                if kind == 'down':
                    x = self._conv(mod, eng.w[id(mod)], x, nchw_in=first, name='down')
                    s *= mod.rate
                    first = False
                elif getattr(mod, 'downsapmle', None) is not None:
                    conv = mod.downsapmle
                    x = self._conv(conv, eng.w[id(conv)], self._block(mod, x, ada), name='down')
                    s *= conv.rate
                elif isinstance(mod, (common.ConvNeXtBlockAdaLN, common.ConvNeXtBlockLN)):
                    x = self._block(mod, x, ada)
                elif isinstance(mod, common.SetKey):
                    feats[mod.key] = x
                else:
                    raise TypeError(f'unsupported encoder module {type(mod)}')
                if qres:
                    feats[H // s] = x
            # ---------------- top-down
            nH, nW = H // m.max_stride, W // m.max_stride
            x = m.bias.reshape(1, 1, 1, -1).expand(B, nH, nW, -1).contiguous()
            kls, zs, li = [], [], 0
            for mod in m.dec_blocks:
                kind = getattr(mod, 'op_kind', None)
                if getattr(mod, 'is_latent_block', False):
                    wl = eng.w[id(mod)]
                    Hs, Ws, zd = x.shape[1], x.shape[2], mod.zdim
                    x = self._block(mod.resnet_front, x, ada)
                    if qres:
                        prior = self._vd(mod.prior, wl['prior'], x)
                        qm = self._vd(mod.posterior, wl['posterior'], x, feats[Hs])
                    else:
                        prior = self._conv(mod.prior, wl['prior'], x, name='prior')
                        e = self._block(mod.posterior0, feats[mod.enc_key], ada)
                        f = self._block(mod.posterior1, x, ada)
                        mg = self._conv(mod.post_merge, wl['post_merge'], f, e, name='post_merge')
                        mg = self._block(mod.posterior2, mg, ada)
                        qm = self._conv(mod.posterior, wl['posterior'], mg, name='posterior')
                    if noise is not None:
                        nz = noise[li].to(eng.device).permute(0, 2, 3, 1).contiguous()
                    else:
                        nz = torch.empty_like(qm).uniform_(-0.5, 0.5)       # same draw order as the reference: one per layer
                    z, kl = _LatentFn.apply(self, qm, prior, nz)
                    kls.append(kl)
                    zs.append(z)
                    li += 1
                    if qres:
                        t = self._conv(mod.z_proj[0], wl['z_proj0'], z, epi=N.EPI_BIAS_GELU, gelu=1, pad_c=wl['z_pad'], name='z_proj0')
                        x = self._conv(mod.z_proj[2], wl['z_proj2'], t, res=x, name='z_proj2')
                    else:
                        x = self._conv(mod.z_proj, wl['z_proj'], z, res=x, name='z_proj')
                    x = self._block(mod.resnet_end, x, ada)
                elif isinstance(mod, (common.ConvNeXtBlockAdaLN, common.ConvNeXtBlockLN)):
                    x = self._block(mod, x, ada)
                elif kind == 'up':
                    conv, r = mod[0], mod.rate
                    last = conv.out_channels // (r * r) == 3
                    x = self._conv(conv, eng.w[id(mod)], x, r=r, nchw_out=last, name='up',
                                   epi=N.EPI_SHUFFLE_NCHW if last else N.EPI_SHUFFLE_NHWC)
                elif isinstance(mod, common.CompresionStopFlag):
                    pass
                else:
                    raise TypeError(f'unsupported decoder module {type(mod)}')
        return dict(x_hat=x, kl=kls, z=zs)

    def objective(self, im, lmb, noise=None, stats=True):
        """The loss of `forward()` of both model classes: mean_b( sum_layers kl_b / ndims + lmb_b * mse_b )
        (qarv/model.py:338-346, qresvae/model.py:533-545) plus the logged statistics (stats=False: no host read-back,
        which is what lets GraphedTrainStep capture the step)."""
        res = self.forward(im, lmb, noise)
        B, imC, imH, imW = im.shape
        ndims = float(imC * imH * imW)
        kl = sum(k.reshape(B, -1).sum(1) for k in res['kl']) / ndims
        x_hat = res['x_hat']
        target = im.sub(0.5).mul_(2.0)
        if getattr(self.model, 'lossless', False):
            # GaussianNLLOutputNet.forward_loss (qresvae/model.py:24-40): the two shuffle heads run on the tcgen05 GEMM (native
            # backward, conv_backward); the per-pixel likelihood is a handful of ATen elementwise ops on [B,3,H,W]
            on, eng = self.model.out_net, self.eng
            heads = [self._conv(h[0], eng.w[id(h)], x_hat, r=h.rate, nchw_out=True, name='up', epi=N.EPI_SHUFFLE_NCHW)
                     for h in (on.conv_mean, on.conv_scale)]
            x_hat = heads[0]
            scale = torch.exp(F.softplus(heads[1] + 16) - 16)
            dist = torch.distributions.Normal(x_hat, scale)
            mass = dist.cdf(target + 0.5 * on.bin_size) - dist.cdf(target - 0.5 * on.bin_size)
            log_prob = torch.where(mass > 1e-6, torch.log(mass.clamp(min=1e-8)), dist.log_prob(target) + math.log(on.bin_size))
            distortion = -log_prob.mean(dim=(1, 2, 3))
        else:
            distortion = (x_hat - target).square().mean(dim=(1, 2, 3))
        loss = (kl + lmb * distortion).mean(0)
        res['loss'] = loss
        if stats == 'device':       # the logged statistics as one device vector (no host read-back: capturable)
            with torch.no_grad():
                im_hat = x_hat.detach().clamp(-1.0, 1.0).mul_(0.5).add_(0.5)
                res['stats_dev'] = torch.stack([kl.mean(0), distortion.mean(0), (im_hat - im).square().mean(),
                                                (lmb * distortion).mean(0)])
                res['kl_layers_dev'] = torch.stack([k.detach().reshape(B, -1).sum(1).mean(0) for k in res['kl']])
            return res
        if not stats:
            return res
        with torch.no_grad():
            im_hat = x_hat.detach().clamp(-1.0, 1.0).mul_(0.5).add_(0.5)
            host = torch.stack([kl.mean(0), distortion.mean(0), (im_hat - im).square().mean(),
                                (lmb * distortion).mean(0)]).cpu()
        res.update(loss=loss, kl_mean=float(host[0]), mse=float(host[1]), im_mse=float(host[2]),
                   lmb_mse=float(host[3]), im_hat=im_hat, kl_img=kl.detach())
        return res


class _GraphedStepFn(torch.autograd.Function):
    """forward = replay of the captured forward graph, backward = replay of the captured backward graph (AutoGraphedTrain).  The
    parameters are inputs of the Function so that autograd routes the captured gradients to their accumulators; the gradient
    tensors handed back are the graph's static buffers themselves (also referenced by the owner, so AccumulateGrad copies
    instead of stealing them: `p.grad` never aliases memory the next replay overwrites)."""

    @staticmethod
    def forward(ctx, ag, im, lmb, *params):
        ag.s_im.copy_(im, non_blocking=True)
        ag.s_lmb.copy_(lmb, non_blocking=True)
        ag.fwd.replay()
        ag.generation += 1                 # the backward graph differentiates the LAST replayed forward
        ctx.ag, ctx.generation = ag, ag.generation
        ctx.mark_non_differentiable(ag.s_stats, ag.s_kl_layers)
        return ag.s_loss.detach().clone(), ag.s_stats, ag.s_kl_layers

    @staticmethod
    def backward(ctx, g_loss, *_):
        ag = ctx.ag
        if ctx.generation != ag.generation or ag.bwd is None:
            raise RuntimeError('this loss belongs to an earlier forward: the graph-replayed training path keeps the activations of the '
                               'latest model(batch) call only (one forward, then its backward).  Set LVAE_TRAIN_AUTOGRAPH=0 or '
                               'model.train_path.autograph_enabled = False for loops that interleave several forward passes.')
        ag.s_gloss.copy_(g_loss)
        ag.bwd.replay()
        return (None, None, None) + tuple(ag.s_grads)


class AutoGraphedTrain:
    """`loss = model(batch)['loss']; loss.backward()` of an UNMODIFIED training loop (lvae/trainer.py:255-300) as two CUDA-graph
    replays: the forward of the first training shape a model sees is captured -- one graph
    for the forward (weight re-packing, all launches, the loss), one for the backward (the native backward kernels with their
    side streams) -- and `model.forward()` in training mode replays them.  The eager path stays for everything else: other
    shapes, supplied noise, `return_rec`, autocast.  Under DistributedDataParallel (lvae/trainer.py:196-203) it works as well:
    the captured gradients reach the REAL parameters' accumulators through _GraphedStepFn, so DDP's reducer hooks fire and its
    bucketed all-reduce runs as usual (after the replay, not overlapped with it): 267 -> 529 images/s on 2 GPUs, parameters
    identical across ranks (scripts/ddp_autograph_check.py, profiles/r2_ddp_autograph.log).  What it removes is the host time of ~1900 Python-issued launches per step: the
    reference-shaped loop is launch-bound (119 ms per step at 16 x 256^2 against 40 ms of GPU work): 135 -> 345 images/s.  On by default
    (LVAE_TRAIN_AUTOGRAPH=0 or model.train_path.autograph_enabled = False: always eager)."""

    def __init__(self, T):
        self.T = T
        self.shape = None
        self.core = None
        self.plan = None            # keeps the scratch buffers the graphs point into alive
        self.generation = 0
        self.fwd = self.bwd = None

    def _signature(self):
        """What the captured graphs depend on besides the batch: the precision mode and every trainable parameter's storage
        address (load_state_dict(assign=True), .to(), freezing a layer change it -> the graphs are dropped and captured again)."""
        m = self.T.model
        return (m.precision, tuple(p.data_ptr() for p in m.parameters() if p.requires_grad))

    def usable(self, im, lmb):
        if torch.is_autocast_enabled() or not im.is_cuda or im.dtype != torch.float32:
            return False
        if (torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1
                and os.environ.get('LVAE_TRAIN_AUTOGRAPH_DDP', '1') == '0'):
            return False
        if self.core is not None and self._signature() != self.sig:
            self.core = self.shape = self.plan = None         # stale: capture again for this batch's shape
            self.fwd = self.bwd = self.s_grads = None
        return self.shape is None or tuple(im.shape) == self.shape

    def __call__(self, im, lmb):
        if self.core is None:
            self._capture(im, lmb)
        return _GraphedStepFn.apply(self, im, lmb, *self.params)

    def _capture(self, im, lmb):
        T = self.T
        dev = T.eng.device
        T.force_refresh = True                      # the captured forward must re-pack the weights the optimizer has just updated
        self.params = [p for p in T.model.parameters() if p.requires_grad]
        keep, T.P = T.P, None                       # a plan of its own: the graphs keep raw pointers into its scratch buffers
        # Capture against ALIAS leaves of the parameters (same storage, fresh autograd identity).  A parameter that has been
        # through an eager backward() owns a gradient accumulator bound to the stream of that step -- normally the legacy default
        # stream -- and it stays alive as long as the caller holds the previous loss; routing a captured backward into it makes
        # the default stream wait on a captured event, which invalidates the capture.  The aliases exist only in here; the graphs
        # read the weights through the same addresses, and _GraphedStepFn hands the captured gradients to the real parameters.
        alias, swapped = {}, []
        for mod in T.model.modules():
            for name, q in list(mod._parameters.items()):
                if q is not None and q.requires_grad:
                    if id(q) not in alias:
                        alias[id(q)] = q.detach().requires_grad_(True)
                    swapped.append((mod, name, q))
                    mod._parameters[name] = alias[id(q)]
        cap_params = [alias[id(q)] for q in self.params]
        try:
            self._capture_graphs(T, dev, im, lmb, cap_params)
        finally:
            for mod, name, q in swapped:
                mod._parameters[name] = q
        self.plan, T.P = T.P, keep
        self.shape = tuple(im.shape)
        self.sig = self._signature()
        self.core = True

    def _capture_graphs(self, T, dev, im, lmb, cap_params):
        with torch.cuda.device(dev):
            self.s_im, self.s_lmb = im.clone(), lmb.clone()
            run = lambda: T.objective(self.s_im, self.s_lmb, stats='device')
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):           # warm-up off the default stream: lazy allocations, autograd stream bookkeeping
                for _ in range(2):
                    res = run()
                    torch.autograd.grad(res['loss'], cap_params, allow_unused=True)
                del res
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            self.fwd, self.bwd = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.fwd):
                res = run()
            self.s_loss, self.s_stats, self.s_kl_layers = res['loss'], res['stats_dev'], res['kl_layers_dev']
            self.s_gloss = torch.ones_like(self.s_loss)
            with torch.cuda.graph(self.bwd, pool=self.fwd.pool()):
                self.s_grads = list(torch.autograd.grad(self.s_loss, cap_params, grad_outputs=self.s_gloss, allow_unused=True))
            del res


def gpu_random_crop_flip(src_u8, crop, hflip=True, generator=None, out=None):
    """The reference's training transform -- RandomCrop(crop) + RandomHorizontalFlip(0.5) + ToTensor
    (lvae/datasets/image.py:45-56, 'crop=256,hflip=True' of train-var-rate.py) -- on the GPU: src_u8 [B,3,Hs,Ws] uint8 on
    the device (decoded images, e.g. a pinned-host batch copied once) -> [B,3,crop,crop] fp32 in [0,1].  Crop origins and
    flip flags are drawn on the device (torch.randint with `generator`), one launch does the rest: no per-sample PIL work,
    no float32 H2D copy (1 byte per element crosses PCIe instead of 4)."""
    assert src_u8.is_cuda and src_u8.dtype == torch.uint8 and src_u8.dim() == 4 and src_u8.shape[1] == 3
    src_u8 = src_u8.contiguous()
    B, _, Hs, Ws = src_u8.shape
    dev = src_u8.device
    y0 = torch.randint(0, Hs - crop + 1, (B,), device=dev, generator=generator, dtype=torch.int32)
    x0 = torch.randint(0, Ws - crop + 1, (B,), device=dev, generator=generator, dtype=torch.int32)
    flip = (torch.rand(B, device=dev, generator=generator) < 0.5).to(torch.uint8) if hflip else torch.zeros(B, device=dev, dtype=torch.uint8)
    out = torch.empty(B, 3, crop, crop, device=dev) if out is None else out
    N.check(N.lib().lvae_crop_flip_u8(src_u8.data_ptr(), y0.data_ptr(), x0.data_ptr(), flip.data_ptr(), out.data_ptr(),
                                      B, Hs, Ws, crop, torch.cuda.current_stream(dev).cuda_stream), 'crop_flip')
    N.launch_count += 1
    return out, (y0, x0, flip)


def allreduce_flat_gradients(params, group, world):
    """Average the gradients of `params` over `group` with ONE all-reduce: the gradients are concatenated into a flat
    buffer, summed over the ranks (NCCL over NVLink / NVSwitch on the GPUs; gloo in the CPU test), scaled by 1 / world and
    copied back in place.  Parameters without a gradient are skipped -- every rank must skip the same ones.
    (Stand-alone helper for eager loops.  GraphedTrainStep does not use it: there the gradients ARE views of one flat
    buffer, reduced bucket by bucket while the backward is still running -- GradientBuckets.)"""
    import torch.distributed as dist
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, group=group)
    flat.mul_(1.0 / world)
    torch._foreach_copy_(grads, [t.view_as(g) for t, g in zip(flat.split([g.numel() for g in grads]), grads)])


class FlatLayout:
    """Offsets of a parameter list inside one flat fp32 buffer; every tensor starts on a 256-byte boundary (TMA maps and
    128-bit loads of the kernels keep working on the views), padding elements stay zero."""
    ALIGN = 64      # elements

    def __init__(self, params):
        self.params = list(params)
        self.offsets, off = [], 0
        for p in self.params:
            self.offsets.append(off)
            off += -(-p.numel() // self.ALIGN) * self.ALIGN
        self.total = off

    def new(self, device, like=None):
        """a zeroed flat buffer; like: list of tensors (same structure as params) whose values are copied in"""
        flat = torch.zeros(self.total, dtype=torch.float32, device=device)
        if like is not None:
            torch._foreach_copy_(self.views(flat), [t.detach() for t in like])
        return flat

    def views(self, flat):
        return [flat[o:o + p.numel()].view(p.shape) for o, p in zip(self.offsets, self.params)]


class GradientBuckets:
    """Data-parallel gradient averaging overlapped with the backward pass (the one collective of the training step:
    reference lvae/trainer.py:63-64 wraps the model in DistributedDataParallel for this).

    The gradients of all parameters are views of ONE flat buffer (FlatLayout), so there is nothing to gather or copy back.
    The buffer is cut into contiguous buckets; a post-accumulate hook per parameter counts arrivals and, when a bucket is
    complete, issues its all-reduce asynchronously (NCCL: ReduceOp.AVG on the communicator's own stream, ordered after the
    kernels that produced the bucket) while autograd keeps running the rest of the backward on the compute stream.
    finish() reduces whatever never completed (parameters without a gradient in this step: every rank sees the same set)
    and makes the compute stream wait for all of them.  Works under CUDA-graph capture (the collectives become graph
    nodes on a forked branch) and on gloo / CPU (SUM + scale), which is how tests/test_sharding_gloo.py covers it."""

    def __init__(self, layout, flat_grad, group, world, bucket_bytes=32 << 20):
        import torch.distributed as dist
        self.dist, self.group, self.world, self.flat = dist, group, world, flat_grad
        self.avg = flat_grad.is_cuda                               # gloo has no AVG
        self.ranges, self.bucket_of, lo, n = [], {}, 0, 0
        for p, off in zip(layout.params, layout.offsets):
            self.bucket_of[id(p)] = len(self.ranges)
            n += 1
            end = off + -(-p.numel() // layout.ALIGN) * layout.ALIGN
            if (end - lo) * 4 >= bucket_bytes:
                self.ranges.append((lo, end, n))
                lo, n = end, 0
        if n:
            self.ranges.append((lo, layout.total, n))
        self.active, self.count, self.fired, self.works = False, [], [], []
        self.hooks = [p.register_post_accumulate_grad_hook(self._hook) for p in layout.params]

    def start(self):
        self.count = [0] * len(self.ranges)
        self.fired = [False] * len(self.ranges)
        self.works, self.active = [], True

    def _hook(self, p):
        if not self.active:
            return
        bi = self.bucket_of[id(p)]
        self.count[bi] += 1
        if self.count[bi] == self.ranges[bi][2] and not self.fired[bi]:
            self._fire(bi)

    def _fire(self, bi):
        lo, hi, _ = self.ranges[bi]
        self.fired[bi] = True
        op = self.dist.ReduceOp.AVG if self.avg else self.dist.ReduceOp.SUM
        self.works.append(self.dist.all_reduce(self.flat[lo:hi], op=op, group=self.group, async_op=True))

    def finish(self):
        for bi in range(len(self.ranges)):
            if not self.fired[bi]:
                self._fire(bi)
        for w in self.works:
            w.wait()
        if not self.avg:
            self.flat.mul_(1.0 / self.world)
        self.active, self.works = False, []

    def remove(self):
        for h in self.hooks:
            h.remove()
        self.hooks = []


class GraphedTrainStep:
    """One training step -- lambda draw, forward, backward, gradient averaging, clipping, optimizer update, EMA, and the
    re-packing of the updated weights into tensor-core operand planes -- captured once into a CUDA graph and replayed per
    batch: the training counterpart of the inference launch plans (the eager step is bound by ~4000 host-side launches).
    Batch shape and lambda policy are fixed per instance.

        step = GraphedTrainStep(model, optimizer, (16, 3, 256, 256), grad_clip=2.0, ema=ema.module)
        loss = step(batch)            # 0-d device tensor, overwritten by the next call

    Memory layout: parameters, gradients, Adam moments and the EMA copy each become views of ONE flat fp32 buffer
    (FlatLayout; `p.data` / `p.grad` are re-pointed, values preserved), so that
      * the gradient all-reduce works on the buffer itself, bucket by bucket, overlapped with the backward
        (GradientBuckets) -- no concatenation, no copy back;
      * clipping + Adam + EMA are two kernel launches over the flat buffers (csrc/optim.cu, `native_optimizer`), instead
        of ~10 torch foreach passes over 907 tensors.
    `optimizer` must be a torch.optim.Adam whose groups share lr / betas / eps with weight_decay 0 (what
    lvae/trainer.py:176-216 builds) for the native update; anything else (or native_optimizer=False) keeps
    `optimizer.step()` inside the graph (the optimizer must then be capturable).  The learning rate and the EMA decay are
    device scalars: set_lr() / set_ema_decay() between replays run the schedule of trainer.py:231-252,373-377.
    """

    def __init__(self, model, optimizer, batch_shape, warmup=1, process_group=None, grad_clip=None, ema=None, ema_decay=0.9999,
                 native_optimizer=None, bucket_bytes=32 << 20):
        """grad_clip: max global gradient norm (clip_grad_norm_ on the reduced gradients, lvae/trainer.py:395);
        ema: a module with the model's parameter structure (timm's ModelEmaV2(model).module or copy.deepcopy(model)) whose
        parameters follow ema += (p - ema)(1 - decay) after every update (trainer.py:374-377).
        process_group: a torch.distributed group (NCCL), or True for the default one -> data parallel (parameters are
        broadcast from rank 0 first, as DDP does).
        warmup: eager steps run before capture (lazy state must exist); parameters, optimizer state and EMA are RESTORED
        afterwards, so the first replay is the first update -- one update per batch, as in the reference's loop."""
        self.model, self.opt = model, optimizer
        self.tp = model.train_path
        dev = model._device()
        self.im = torch.zeros(batch_shape, device=dev)
        self.grad_clip, self.ema = grad_clip, ema
        self.pg, self.world = None, 1
        if process_group is not None:
            import torch.distributed as dist
            self.pg = dist.group.WORLD if process_group is True else process_group
            self.world = dist.get_world_size(self.pg)
            for p in model.parameters():                         # same starting point on every rank (DDP does the same)
                dist.broadcast(p.data, src=dist.get_global_rank(self.pg, 0), group=self.pg)
        # ---- flat storage
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.layout = FlatLayout(self.params)
        self.flat_p = self.layout.new(dev, like=self.params)
        self.flat_g = self.layout.new(dev)
        for p, vp, vg in zip(self.params, self.layout.views(self.flat_p), self.layout.views(self.flat_g)):
            p.data = vp
            p.grad = vg
        model.engine.invalidate()
        self.flat_e = None
        if ema is not None:
            eps_ = [q for q, p in zip(ema.parameters(), model.parameters()) if p.requires_grad]
            self.flat_e = self.layout.new(dev, like=eps_)
            for q, ve in zip(eps_, self.layout.views(self.flat_e)):
                q.data = ve
        self.native = self._native_ok(optimizer) if native_optimizer is None else bool(native_optimizer)
        if self.native and not self._native_ok(optimizer):
            raise ValueError('native_optimizer needs torch.optim.Adam with weight_decay 0, amsgrad False and one (lr, betas, eps) for all groups')
        g0 = optimizer.param_groups[0]
        self.lr = torch.tensor(float(g0['lr']), device=dev)
        self.ema_decay = torch.tensor([float(ema_decay), 1.0 - float(ema_decay)], device=dev)      # (decay, 1 - decay)
        self.step_t = torch.zeros((), device=dev)
        self.grad_norm = torch.zeros((), device=dev)             # global gradient norm of the last step (before clipping)
        if self.native:
            self.betas, self.eps = tuple(g0['betas']), float(g0['eps'])
            self.flat_m, self.flat_v = self.layout.new(dev), self.layout.new(dev)
            self.scratch = torch.zeros(N.lib().lvae_optim_scratch_doubles(), dtype=torch.float64, device=dev)
        self.buckets = GradientBuckets(self.layout, self.flat_g, self.pg, self.world, bucket_bytes) if self.world > 1 else None
        self.graph, self.loss, self.warmup = None, None, max(1, warmup)   # >= 1: lazy state (packed weights, scratch) exists before capture
        # the gradient views were created on the current stream, warm-up and capture run on side streams: intended
        if hasattr(torch.autograd.graph, 'set_warn_on_accumulate_grad_stream_mismatch'):
            torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)

    @staticmethod
    def _native_ok(opt):
        if type(opt) is not torch.optim.Adam:
            return False
        gs = opt.param_groups
        same = all((g['lr'], tuple(g['betas']), g['eps']) == (gs[0]['lr'], tuple(gs[0]['betas']), gs[0]['eps']) for g in gs)
        return same and all(g['weight_decay'] == 0 and not g['amsgrad'] and not g.get('maximize', False) for g in gs) and not opt.state

    def set_lr(self, lr):
        self.lr.fill_(float(lr))
        if not self.native:
            for g in self.opt.param_groups:
                if isinstance(g['lr'], torch.Tensor):
                    g['lr'].fill_(float(lr))
                else:
                    g['lr'] = float(lr)      # takes effect at the next capture only

    def set_ema_decay(self, decay):
        self.ema_decay.copy_(torch.tensor([float(decay), 1.0 - float(decay)]))

    def _step(self):
        m, B = self.model, self.im.shape[0]
        lmb = m._lmb(B) if self.tp.family == 'qres' else m.sample_lmb(B)
        self.tp.force_refresh = True
        try:
            res = self.tp.objective(self.im, lmb, stats=False)
        finally:
            self.tp.force_refresh = False
        self.flat_g.zero_()                                       # the gradients are views of it: autograd accumulates in place
        if self.buckets is not None:
            self.buckets.start()
        res['loss'].backward()
        if self.buckets is not None:
            self.buckets.finish()
        if self.native:
            self.step_t.add_(1.0)
            N.check(N.lib().lvae_adam_clip_ema(
                self.flat_p.data_ptr(), self.flat_g.data_ptr(), self.flat_m.data_ptr(), self.flat_v.data_ptr(),
                0 if self.flat_e is None else self.flat_e.data_ptr(), self.layout.total, self.scratch.data_ptr(),
                float(self.grad_clip or 0.0), self.lr.data_ptr(), self.step_t.data_ptr(), self.ema_decay.data_ptr(),
                self.betas[0], self.betas[1], self.eps, self.grad_norm.data_ptr(), self.tp.eng._stream()), 'adam_clip_ema')
            N.launch_count += 2
        else:
            if self.grad_clip is not None:
                self.grad_norm.copy_(torch.nn.utils.clip_grad_norm_(self.params, self.grad_clip, foreach=True))
            self.opt.step()
            if self.flat_e is not None:
                self.flat_e.mul_(self.ema_decay[0]).add_(self.flat_p * self.ema_decay[1])
        return res['loss'].detach()

    def _snapshot(self):
        st = dict(p=self.flat_p.clone(), e=None if self.flat_e is None else self.flat_e.clone(), t=self.step_t.clone())
        if self.native:
            st['m'], st['v'] = self.flat_m.clone(), self.flat_v.clone()
        else:
            import copy
            st['opt'] = copy.deepcopy(self.opt.state_dict())
        return st

    def _restore(self, st):
        self.flat_p.copy_(st['p'])
        self.step_t.copy_(st['t'])
        if st['e'] is not None:
            self.flat_e.copy_(st['e'])
        if self.native:
            self.flat_m.copy_(st['m'])
            self.flat_v.copy_(st['v'])
        else:
            # in place: the captured graph must keep pointing at the state tensors the warm-up created
            cur = self.opt.state_dict()['state']
            for k, d in st['opt']['state'].items():
                for name, val in d.items():
                    if isinstance(val, torch.Tensor):
                        cur[k][name].copy_(val)
            if not st['opt']['state']:
                for d in cur.values():
                    for val in d.values():
                        if isinstance(val, torch.Tensor):
                            val.zero_()

    def __call__(self, im):
        assert self.model.training and tuple(im.shape) == tuple(self.im.shape)
        self.im.copy_(im, non_blocking=True)
        if self.graph is None:
            dev = self.im.device
            snap = self._snapshot()
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(self.warmup):
                    self._step()
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            self.graph = torch.cuda.CUDAGraph()
            n0 = N.launch_count
            with torch.cuda.graph(self.graph):
                self.loss = self._step()
            self.launches_per_replay = N.launch_count - n0        # liblvae_b200 launches recorded in the graph
            N.launch_count = n0
            self._restore(snap)                                   # the warm-up steps leave no trace: one update per batch
            del snap
        self.graph.replay()
        N.launch_count += self.launches_per_replay
        self.tp.eng.invalidate()        # the replay updated the parameters in place: the next eager / inference use re-packs
        return self.loss

    def release(self):
        """Drop the captured graph (it holds the recorded NCCL work and every buffer of the step).  Call before
        torch.distributed.destroy_process_group(): a live graph with collectives in it keeps the communicator busy."""
        dev = self.im.device
        torch.cuda.synchronize(dev)
        self.graph, self.loss = None, None
        if self.buckets is not None:
            self.buckets.remove()
        torch.cuda.synchronize(dev)
