"""TransMIL aggregator on the B200 primitives -- inference.

Mirrors ``TransMIL`` (src/stamp/modeling/models/trans_mil.py:276-326; ``ModelName.TRANS_MIL``,
src/stamp/modeling/registry.py:52-58): same constructor ``(dim_output, dim_input, dim_hidden)``, same sub-modules and
state-dict keys (``_fc1.0``, ``cls_token``, ``layer1/2.norm``, ``layer1/2.attn.to_qkv``, ``.to_out.0``,
``.res_conv``, ``pos_layer.proj/proj1/proj2``, ``norm``, ``_fc2``), ``forward(h [B, n, dim_input]) -> logits [B, C]``.

  _fc1 Linear + ReLU (the large product)        stamp_gemm_tn (tcgen05, fp16 operands, fp32 out)
  LayerNorm                                     stamp_layernorm (fp32 out)
  Nystrom attention (:93-160)                   fp32 on the CUDA cores (csrc/transmil.cu): to_qkv / to_out, landmark means, softmax of
                                                q_l k_l^T, Newton-Schulz pseudo-inverse (6 iterations), attn3 @ v and
                                                attn1 @ (pinv @ (attn3 @ v)) as two fp32 attention calls, depth-wise
                                                residual convolution over the tokens
  PPEG (:253-273)                               the 7x7, 5x5, 3x3 depth-wise kernels and the identity folded into ONE 7x7
                                                depth-wise convolution (stamp_dwconv2d_f32)

``(attn1 @ attn2_inv) @ (attn3 @ v)`` is evaluated as ``attn1 @ (attn2_inv @ (attn3 @ v))`` (associativity: the
n x 256 product with the pseudo-inverse becomes a 256 x 64 one).  The bags of a batch share the scaling of the
pseudo-inverse's start, as in the reference (see ``_nystrom_layer``).  Measured against the reference module: 6e-4 per
bag, 1e-6 with ``fc1_fp32 = True``.  Inference only; dropout is the identity in eval mode.
"""

from __future__ import annotations

import ctypes as C
import math

import torch
from torch import Tensor, nn

from . import _lib, ops


def _bind() -> C.CDLL:
    lib = _lib.load()
    if not getattr(lib, "_transmil_bound", False):
        ll, vp, i, f = C.c_longlong, C.c_void_p, C.c_int, C.c_float
        lib.stamp_landmark_mean_f32.argtypes = [vp, ll, i, vp, ll, i, i, vp]
        lib.stamp_sgemm_batched_f32.argtypes = [vp, ll, ll, vp, ll, ll, vp, ll, ll, i, i, i, i, i, f, f, vp, i, vp]
        lib.stamp_softmax_rows_f32.argtypes = [vp, ll, ll, i, i, i, vp]
        lib.stamp_pinv_init_f32.argtypes = [vp, vp, i, i, vp, vp]
        lib.stamp_attention_f32.argtypes = [vp, ll, vp, ll, vp, ll, vp, ll, i, i, i, f, vp, i, vp]
        lib.stamp_dwconv1d_add_f32.argtypes = [vp, ll, vp, vp, ll, i, i, i, vp]
        lib.stamp_dwconv2d_f32.argtypes = [vp, ll, vp, vp, vp, ll, i, i, i, i, vp]
        for name in ("stamp_landmark_mean_f32", "stamp_sgemm_batched_f32", "stamp_softmax_rows_f32", "stamp_pinv_init_f32",
                     "stamp_attention_f32", "stamp_dwconv1d_add_f32", "stamp_dwconv2d_f32"):
            getattr(lib, name).restype = C.c_int
        lib._transmil_bound = True
    return lib


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class _NystromAttention(nn.Module):
    """Parameter container with the reference's names (trans_mil.py:48-91)."""

    def __init__(self, dim: int, dim_head: int, heads: int, num_landmarks: int, pinv_iterations: int = 6,
                 residual_conv_kernel: int = 33, dropout: float = 0.1) -> None:
        super().__init__()
        self.heads, self.dim_head, self.num_landmarks, self.pinv_iterations = heads, dim_head, num_landmarks, pinv_iterations
        inner = heads * dim_head
        self.to_qkv = nn.Linear(dim, inner * 3, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner, dim), nn.Dropout(dropout))
        self.res_conv = nn.Conv2d(heads, heads, (residual_conv_kernel, 1), padding=(residual_conv_kernel // 2, 0),
                                  groups=heads, bias=False)


class _Transformer(nn.Module):
    def __init__(self, dim: int) -> None:
        super().__init__()
        self.norm = nn.LayerNorm(dim)
        self.attn = _NystromAttention(dim=dim, dim_head=dim // 8, heads=8, num_landmarks=dim // 2)


class _PPEG(nn.Module):
    def __init__(self, dim: int) -> None:
        super().__init__()
        self.proj = nn.Conv2d(dim, dim, 7, 1, 3, groups=dim)
        self.proj1 = nn.Conv2d(dim, dim, 5, 1, 2, groups=dim)
        self.proj2 = nn.Conv2d(dim, dim, 3, 1, 1, groups=dim)


class TransMIL(nn.Module):
    def __init__(self, dim_output: int, dim_input: int, dim_hidden: int) -> None:
        super().__init__()
        if dim_hidden % 8 or dim_hidden // 8 != 64:
            raise ValueError("unsupported TransMIL configuration for the sm_100a kernels (dim_hidden 512 = 8 heads of 64; "
                             "any dim_input: widths that are not a multiple of 8 run zero-padded)")
        self.pos_layer = _PPEG(dim_hidden)
        self._fc1 = nn.Sequential(nn.Linear(dim_input, dim_hidden), nn.ReLU())
        self.cls_token = nn.Parameter(torch.randn(1, 1, dim_hidden))
        self.n_classes = dim_output
        self.layer1 = _Transformer(dim_hidden)
        self.layer2 = _Transformer(dim_hidden)
        self.norm = nn.LayerNorm(dim_hidden)
        self._fc2 = nn.Linear(dim_hidden, dim_output)
        self._half: dict[int, tuple[tuple, Tensor]] = {}
        self.fc1_fp32 = False   # True: _fc1 on the fp32 path too (1e-6 instead of 6e-4 against the reference module)

    def _w16(self, p: Tensor) -> Tensor:
        """fp16 copy of a weight matrix (columns zero-padded to a multiple of 8, the GEMM's K granularity), refreshed when
        the parameter changes."""
        try:
            key = (p._version, p.data_ptr())
        except RuntimeError:            # parameters created under inference_mode do not track versions
            key = (-1, p.data_ptr())
        hit = self._half.get(id(p))
        if hit is None or hit[0] != key or hit[1].device != p.device:
            w = p.detach().half()
            if w.shape[1] % 8:
                w = torch.nn.functional.pad(w, (0, 8 - w.shape[1] % 8))
            hit = (key, w.contiguous())
            self._half[id(p)] = hit
        return hit[1]

    # ---- one Transformer layer for all bags of the batch: x_b += NystromAttention(LayerNorm(x_b))  (x_b fp32 [n, C], in place)
    def _nystrom_layer(self, xs: list[Tensor], layer: _Transformer) -> None:
        """The bags of a batch go through the layer together because the reference's pseudo-inverse start
        ``z0 = x^T / (max row sum * max column sum)`` takes its maxima over the WHOLE batch and all heads
        (``torch.max`` of the full tensor, trans_mil.py:31-35): with six Newton-Schulz steps the scaling of the start is
        still visible in the result (1e-3 on the logits), so a bag's output depends on its batch mates -- reproduced."""
        lib = _bind()
        a = layer.attn
        n, Cd = xs[0].shape
        B, H, m, hd = len(xs), a.heads, a.num_landmarks, a.dim_head
        scale = hd ** -0.5
        dev, st = xs[0].device, _stream()
        pad = (m - n % m) % m
        n_p = n + pad
        grp = math.ceil(n / m)                                      # tokens per landmark (n_p == m * grp)
        mm = lib.stamp_sgemm_batched_f32
        wq = a.to_qkv.weight.detach().float().contiguous()
        a2 = torch.empty((B * H, m, m), dtype=torch.float32, device=dev)
        per_bag = []
        for b, x in enumerate(xs):
            y = torch.zeros((n_p, Cd), dtype=torch.float32, device=dev)  # zero rows in FRONT (:104-109)
            ops.layernorm(x, layer.norm.weight, layer.norm.bias, layer.norm.eps, torch.float32, out=y[pad:])
            qkv = torch.empty((n_p, 3 * Cd), dtype=torch.float32, device=dev)
            # to_qkv in fp32: q and k feed the landmark softmaxes and the truncated pseudo-inverse
            _lib.check(mm(y.data_ptr(), Cd, 0, wq.data_ptr(), Cd, 0, qkv.data_ptr(), 3 * Cd, 0, n_p, 3 * Cd, Cd, 1, 1, 1.0, 0.0,
                          None, 0, st), "sgemm")
            ld = qkv.stride(0)
            ql = torch.empty((m, Cd), dtype=torch.float32, device=dev)
            kl = torch.empty((m, Cd), dtype=torch.float32, device=dev)
            for src, dst in ((qkv[:, :Cd], ql), (qkv[:, Cd:2 * Cd], kl)):
                _lib.check(lib.stamp_landmark_mean_f32(src.data_ptr(), ld, grp, dst.data_ptr(), Cd, m, Cd, st), "stamp_landmark_mean_f32")
            # attn2 = softmax(q_l k_l^T * scale) per head
            _lib.check(mm(ql.data_ptr(), Cd, hd, kl.data_ptr(), Cd, hd, a2[b * H:].data_ptr(), m, m * m, m, m, hd, H, 1, scale, 0.0,
                          None, 0, st), "sgemm")
            per_bag.append((qkv, ql, kl))
        _lib.check(lib.stamp_softmax_rows_f32(a2.data_ptr(), m, m * m, m, m, B * H, st), "stamp_softmax_rows_f32")
        # truncated pseudo-inverse of all B * H matrices, one common scaling of the start
        z = torch.empty_like(a2)
        scratch = torch.empty(2, dtype=torch.int32, device=dev)
        _lib.check(lib.stamp_pinv_init_f32(a2.data_ptr(), z.data_ptr(), m, B * H, scratch.data_ptr(), st), "stamp_pinv_init_f32")
        xz, t1, t2 = torch.empty_like(a2), torch.empty_like(a2), torch.empty_like(a2)
        sq = lambda A, Bm, Cm, alpha, eye: _lib.check(mm(A.data_ptr(), m, m * m, Bm.data_ptr(), m, m * m, Cm.data_ptr(), m, m * m,  # noqa: E731
                                                         m, m, m, B * H, 0, alpha, eye, None, 0, st), "sgemm")
        for _ in range(a.pinv_iterations):                          # z <- 1/4 z (13 I - xz (15 I - xz (7 I - xz)))
            sq(a2, z, xz, 1.0, 0.0)
            sq(xz, xz, t1, 1.0, 7.0)
            sq(xz, t1, t2, 1.0, 15.0)
            sq(z, t2, t1, 0.25, 13.0)
            z, t1 = t1, z

        def att(q_, ldq, k_, ldk, v_, ldv, o_, nq, nk):
            """softmax(scale q k^T) v per head; few query blocks over many keys (the landmark rows) share the keys among CTAs"""
            ctas = -(-nq // 64) * H
            splits = max(1, min(-(-148 * 4 // ctas), nk // 128)) if ctas < 148 else 1
            part = torch.empty(splits * H * nq * 66, dtype=torch.float32, device=dev) if splits > 1 else None
            _lib.check(lib.stamp_attention_f32(q_.data_ptr(), ldq, k_.data_ptr(), ldk, v_.data_ptr(), ldv, o_.data_ptr(), Cd, nq, nk, H,
                                               scale, part.data_ptr() if part is not None else None, splits, st), "attention_f32")

        taps = a.res_conv.weight.shape[2]
        wconv = a.res_conv.weight.detach().float().reshape(H, taps).contiguous()
        lin = a.to_out[0]
        wo, bo = lin.weight.detach().float().contiguous(), lin.bias.detach().float().contiguous()
        for b, (x, (qkv, ql, kl)) in enumerate(zip(xs, per_bag)):
            ld = qkv.stride(0)
            q, k, v = qkv[:, :Cd], qkv[:, Cd:2 * Cd], qkv[:, 2 * Cd:]
            # attn3 @ v (landmark queries over all keys), W = pinv @ that, out = attn1 @ W (all queries over the landmarks)
            a3v = torch.empty((m, Cd), dtype=torch.float32, device=dev)
            att(ql, Cd, k, ld, v, ld, a3v, m, n_p)
            w = torch.empty((m, Cd), dtype=torch.float32, device=dev)
            _lib.check(mm(z[b * H:].data_ptr(), m, m * m, a3v.data_ptr(), Cd, hd, w.data_ptr(), Cd, hd, m, hd, m, H, 0, 1.0, 0.0,
                          None, 0, st), "sgemm")
            out = torch.empty((n_p, Cd), dtype=torch.float32, device=dev)
            att(q, ld, kl, Cd, w, Cd, out, n_p, m)
            _lib.check(lib.stamp_dwconv1d_add_f32(v.data_ptr(), ld, wconv.data_ptr(), out.data_ptr(), Cd, n_p, H, taps, st),
                       "stamp_dwconv1d_add_f32")
            # x += to_out(out[-n:])
            _lib.check(mm(out[pad:].data_ptr(), Cd, 0, wo.data_ptr(), Cd, 0, x.data_ptr(), Cd, 0, n, Cd, Cd, 1, 1, 1.0, 0.0,
                          bo.data_ptr(), 1, st), "sgemm")

    def _ppeg(self, x: Tensor, side: int) -> Tensor:
        """x fp32 [1 + side*side, C] -> same shape; class token untouched (:263-273)."""
        lib = _bind()
        p = self.pos_layer
        Cd = x.shape[1]
        k = p.proj.weight.detach().float().reshape(Cd, 7, 7).clone()
        k[:, 1:6, 1:6] += p.proj1.weight.detach().float().reshape(Cd, 5, 5)
        k[:, 2:5, 2:5] += p.proj2.weight.detach().float().reshape(Cd, 3, 3)
        k[:, 3, 3] += 1.0                                           # + cnn_feat
        bias = (p.proj.bias + p.proj1.bias + p.proj2.bias).detach().float().contiguous()
        out = torch.empty_like(x)
        out[0] = x[0]
        _lib.check(lib.stamp_dwconv2d_f32(x[1:].data_ptr(), Cd, k.contiguous().data_ptr(), bias.data_ptr(), out[1:].data_ptr(), Cd,
                                          side, side, Cd, 7, _stream()), "stamp_dwconv2d_f32")
        return out

    def forward(self, h: Tensor, **kwargs) -> Tensor:
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise NotImplementedError("stamp_b200 TransMIL is inference-only: call it under torch.no_grad() / "
                                      "inference_mode() (training runs through the reference module)")
        if not h.is_cuda or not self.cls_token.is_cuda:
            raise RuntimeError("stamp_b200 TransMIL runs on a CUDA device only (no CPU fallback)")
        if self.training:
            raise RuntimeError("call .eval() first: dropout of the training mode is not implemented")
        B, n, F = h.shape
        Cd = self.cls_token.shape[-1]
        fc1 = self._fc1[0]
        side = int(math.ceil(math.sqrt(n)))
        add = side * side - n
        xs = []
        for b in range(B):
            feats = torch.empty((n, Cd), dtype=torch.float32, device=h.device)
            if self.fc1_fp32:
                hb = h[b].detach().float().contiguous()
                w1, b1 = fc1.weight.detach().float().contiguous(), fc1.bias.detach().float().contiguous()
                _lib.check(_bind().stamp_sgemm_batched_f32(hb.data_ptr(), F, 0, w1.data_ptr(), F, 0, feats.data_ptr(), Cd, 0, n, Cd, F,
                                                           1, 1, 1.0, 0.0, b1.data_ptr(), 2, _stream()), "sgemm")
            else:
                hb = h[b].detach().half()
                if F % 8:                                               # zero columns against the zero-padded weight columns
                    hb = torch.nn.functional.pad(hb, (0, 8 - F % 8))
                ops.gemm_tn(hb.contiguous(), self._w16(fc1.weight), out=feats, bias=fc1.bias.detach().float(),
                            act=ops.ACT_RELU, store=ops.ST_32)
            xs.append(torch.cat([self.cls_token.detach().float().reshape(1, Cd), feats, feats[:add]], dim=0).contiguous())
        self._nystrom_layer(xs, self.layer1)
        xs = [self._ppeg(x, side) for x in xs]
        self._nystrom_layer(xs, self.layer2)
        logits = []
        for x in xs:
            cls = torch.nn.functional.layer_norm(x[0], (Cd,), self.norm.weight.float(), self.norm.bias.float(), self.norm.eps)
            logits.append(torch.nn.functional.linear(cls, self._fc2.weight.float(), self._fc2.bias.float()))
        return torch.stack(logits)
