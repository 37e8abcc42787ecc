"""Drop-in MIL backbone: STAMP's ALiBi ``VisionTransformer`` with its forward running in
``libstamp_b200.so``.

Contract mirrored from the reference (src/stamp/modeling/models/vision_tranformer.py:298-384 and
SURVEY.md 8b):

* constructor ``VisionTransformer(*, dim_output, dim_input, dim_model, n_layers, n_heads,
  dim_feedforward, dropout, use_alibi)`` -- ``Base._build_backbone`` filters Lightning kwargs by this
  signature (src/stamp/modeling/models/__init__.py:112-131);
* ``forward(bags [B,N,F], *, coords [B,N,2], mask [B,N] | None) -> [B,C]``;
* identical parameter / buffer names, shapes and initialisation, so ``.ckpt`` files written by the
  reference load here and vice versa (``LitX.load_from_checkpoint(model_class=VisionTransformer)``,
  src/stamp/modeling/deploy.py:49-58).

The sub-modules below only hold parameters under the reference's names; the arithmetic is one call
to ``stamp_mil_forward`` (per-head Q/K/V Linears are packed into one [3d, d] GEMM operand, see
``_pack``) under ``torch.no_grad()`` / ``inference_mode()``; with autograd enabled the call goes
through the checkpointing forward and the backward kernels of ``stamp_b200.train`` (bf16 operands,
mask=None branch, ALiBi or nn.MultiheadAttention variant).
"""

from __future__ import annotations

import ctypes as C

import torch
from torch import Tensor, nn

from . import _lib


class StampMilConfig(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("dim_input", "dim_model", "n_layers", "n_heads", "dim_ff",
                                       "dim_output", "use_alibi", "dim_model_real", "head_dim_real")]


class StampMilWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("proj_w", "proj_b", "class_token", "norm_w", "norm_b",
                                          "head_w", "head_b")]


class StampMilLayer(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("ln1_w", "ln1_b", "qkv_w", "qkv_b", "v_w3", "slope", "fc_w",
                                          "fc_b", "ln2_w", "ln2_b", "ff1_w", "ff1_b", "ff2_w", "ff2_b")]


def _bind() -> C.CDLL:
    lib = _lib.load()
    if not getattr(lib, "_mil_bound", False):
        lib.stamp_mil_workspace_bytes.restype = C.c_size_t
        lib.stamp_mil_workspace_bytes.argtypes = [C.POINTER(StampMilConfig), C.c_int, C.c_int]
        lib.stamp_mil_forward.restype = C.c_int
        lib.stamp_mil_forward.argtypes = [C.POINTER(StampMilConfig), C.POINTER(StampMilWeights),
                                          C.POINTER(StampMilLayer), C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]
        lib.stamp_mil_ragged_workspace_bytes.restype = C.c_size_t
        lib.stamp_mil_ragged_workspace_bytes.argtypes = [C.POINTER(StampMilConfig), C.c_int, C.c_int, C.c_int]
        lib.stamp_mil_forward_ragged.restype = C.c_int
        lib.stamp_mil_forward_ragged.argtypes = [C.POINTER(StampMilConfig), C.POINTER(StampMilWeights),
                                                 C.POINTER(StampMilLayer), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                                 C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        lib._mil_bound = True
    return lib


# Bumped by anything that rewrites parameters through raw device pointers (FusedAdamW.step): the parameters'
# own version counters do not see such writes, so the packed-weight cache is keyed on this as well.
_weights_epoch = 0


def bump_weights_epoch() -> None:
    global _weights_epoch
    _weights_epoch += 1


def round_to_tf32(t: Tensor) -> Tensor:
    """Round-to-nearest onto TF32's 10 explicit mantissa bits (the tensor core would truncate)."""
    i = t.detach().float().contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


# ---- parameter containers named exactly like the reference's modules ---------------------------
class _RunningMeanScaler(nn.Module):
    def __init__(self) -> None:
        super().__init__()
        self.running_mean = nn.Buffer(torch.ones(1))
        self.items_so_far = nn.Buffer(torch.ones(1))


class _ALiBi(nn.Module):
    def __init__(self) -> None:
        super().__init__()
        self.scale_distance = _RunningMeanScaler()
        self.bias_scale = nn.Parameter(torch.rand(1))


class MultiHeadALiBi(nn.Module):
    def __init__(self, *, embed_dim: int, num_heads: int) -> None:
        super().__init__()
        if embed_dim % num_heads != 0:
            raise ValueError(f"{embed_dim=} has to be divisible by {num_heads=}")
        hd = embed_dim // num_heads
        mk = lambda: nn.ModuleList([nn.Linear(embed_dim, hd) for _ in range(num_heads)])
        self.query_encoders, self.key_encoders, self.value_encoders = mk(), mk(), mk()
        self.attentions = nn.ModuleList([_ALiBi() for _ in range(num_heads)])
        self.fc = nn.Linear(embed_dim, embed_dim)


class SelfAttention(nn.Module):
    def __init__(self, *, dim: int, num_heads: int, dropout: float, use_alibi: bool) -> None:
        super().__init__()
        self.heads = num_heads
        self.norm = nn.LayerNorm(dim)
        self.mhsa = (MultiHeadALiBi(embed_dim=dim, num_heads=num_heads) if use_alibi
                     else nn.MultiheadAttention(dim, num_heads, dropout, batch_first=True))


def feed_forward(dim: int, hidden_dim: int, dropout: float = 0.5) -> nn.Sequential:
    return nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, hidden_dim), nn.GELU(), nn.Dropout(dropout),
                         nn.Linear(hidden_dim, dim), nn.Dropout(dropout))


class Transformer(nn.Module):
    def __init__(self, *, dim: int, depth: int, heads: int, mlp_dim: int, dropout: float,
                 use_alibi: bool) -> None:
        super().__init__()
        self.depth = depth
        self.layers = nn.ModuleList([
            nn.ModuleList([SelfAttention(dim=dim, num_heads=heads, dropout=dropout, use_alibi=use_alibi),
                           feed_forward(dim, mlp_dim)])
            for _ in range(depth)])
        self.norm = nn.LayerNorm(dim)


class VisionTransformer(nn.Module):
    def __init__(self, *, dim_output: int, dim_input: int, dim_model: int, n_layers: int, n_heads: int,
                 dim_feedforward: int, dropout: float, use_alibi: bool) -> None:
        super().__init__()
        self.class_token = nn.Parameter(torch.randn(dim_model))
        self.project_features = nn.Sequential(nn.Linear(dim_input, dim_model, bias=True), nn.GELU(),
                                              nn.Dropout(dropout))
        self.transformer = Transformer(dim=dim_model, depth=n_layers, heads=n_heads,
                                       mlp_dim=dim_feedforward, dropout=dropout, use_alibi=use_alibi)
        self.mlp_head = nn.Sequential(nn.Linear(dim_model, dim_output))
        self._cfg = dict(dim_input=dim_input, dim_model=dim_model, n_layers=n_layers, n_heads=n_heads,
                         dim_ff=dim_feedforward, dim_output=dim_output, use_alibi=int(use_alibi))
        self._packed = None       # (key, tensors kept alive, cfg, weights, layers)
        self._key_tensors = None
        self._workspace: dict[int, Tensor] | None = None   # per CUDA stream: concurrent bags do not share scratch
        self._train_ctx: Tensor | None = None   # checkpoints between a training forward and its backward
        self._train_gen = 0
        self._train_state = None
        self._flat_opt = None            # weakref to a FusedAdamW(model=self): packed-order flat buffers (train.py)
        self._grad_sink_views = None

    # ---- weight packing: reference layout -> GEMM operands (cached until a parameter changes) ----
    def _pack_key(self):
        # (this runs on every inference call: one attribute read per tensor; the storage pointers of the first and
        #  the last one catch .to(device) / load_state_dict(assign=True), which replace all of them together)
        ps = self._key_tensors
        if ps is None:
            ps = self._key_tensors = list(self.parameters()) + list(self.buffers())
        return (_weights_epoch, ps[0].device, tuple(p._version for p in ps), ps[0].data_ptr(), ps[-1].data_ptr())

    # ctypes structures with raw device pointers, workspaces and the checkpoint buffer are per-process caches:
    # copy.deepcopy / torch.save(model) / spawn pickling carry the parameters only
    _TRANSIENT = {"_packed": None, "_workspace": None, "_train_ctx": None, "_train_state": None, "_key_tensors": None,
                  "_flat_opt": None, "_grad_sink_views": None}

    def __getstate__(self):
        state = dict(super().__getstate__() if hasattr(super(), "__getstate__") else self.__dict__)
        state.update(self._TRANSIENT)
        return state

    def __setstate__(self, state):
        super().__setstate__(state)
        for k, v in self._TRANSIENT.items():
            self.__dict__.setdefault(k, v)
        self.__dict__.setdefault("_train_gen", 0)

    # ---- shapes outside the kernels' envelope run zero-padded (include/stamp_b200.h, StampMilConfig) ----
    def _shape_plan(self) -> dict:
        """Kernel shapes: widths in multiples of 8, heads of 32 / 64 (/ 80 without ALiBi) columns.  Anything else --
        the reference's unit tests use heads of 33 and 34, 457 input features, 135 hidden units
        (tests/test_model.py) -- is zero-padded: ``dim_model`` becomes ``n_heads * padded head width``."""
        c = self._cfg
        d, H, F, ff = c["dim_model"], c["n_heads"], c["dim_input"], c["dim_ff"]
        hd = d // H
        if hd <= 32:
            hp = 32
        elif hd <= 64:
            hp = 64
        elif hd <= 80 and not c["use_alibi"]:
            hp = 80
        else:
            raise ValueError(f"unsupported MIL configuration for the sm_100a kernels: head dimension {hd} "
                             f"(at most 64 with ALiBi, 80 without)")
        up8 = lambda n: (n + 7) // 8 * 8
        plan = dict(d=d, hd=hd, hp=hp, dc=H * hp, F=F, Fp=up8(F), ff=ff, ffp=up8(ff))
        plan["padded"] = plan["dc"] != d or plan["Fp"] != F or plan["ffp"] != ff
        return plan

    def _pack(self):
        key = self._pack_key()
        if self._packed is not None and self._packed[0] == key:
            return self._packed
        plan = self._shape_plan()
        if plan["padded"]:
            self._packed = self._pack_padded(key, plan)
            return self._packed
        keep: list[Tensor] = []

        def f32(t: Tensor) -> int:
            t = t.detach().float().contiguous()
            keep.append(t)
            return t.data_ptr()

        def f16(t: Tensor) -> int:
            t = t.detach().half().contiguous()
            keep.append(t)
            return t.data_ptr()

        cfg = StampMilConfig(**self._cfg)
        use_alibi = bool(self._cfg["use_alibi"])
        w = StampMilWeights(f16(self.project_features[0].weight), f32(self.project_features[0].bias),
                            f32(self.class_token), f32(self.transformer.norm.weight),
                            f32(self.transformer.norm.bias), f32(self.mlp_head[0].weight),
                            f32(self.mlp_head[0].bias))
        layers = (StampMilLayer * max(1, self._cfg["n_layers"]))()
        for i, (att, ff) in enumerate(self.transformer.layers):
            m = att.mhsa
            if use_alibi:
                qkv_w = torch.cat([e.weight for grp in (m.query_encoders, m.key_encoders, m.value_encoders) for e in grp])
                qkv_b = torch.cat([e.bias for grp in (m.query_encoders, m.key_encoders, m.value_encoders) for e in grp])
                slope = torch.cat([a.bias_scale / a.scale_distance.running_mean for a in m.attentions])
                # split precision (hi + lo) for the two contractions the ALiBi term flows through,
                # K-concatenated so each is ONE GEMM: [hi | lo] read as hi, lo, hi against [W_hi | W_hi | W_lo]
                fc_full = m.fc.weight.detach().float()
                fc_hi = round_to_tf32(fc_full)
                fc_lo = round_to_tf32(fc_full - fc_hi)
                fc_w, fc_b, slope_p = f32(torch.cat([fc_hi, fc_hi, fc_lo], dim=1)), f32(m.fc.bias), f32(slope)
                wv = qkv_w[2 * self._cfg["dim_model"]:].detach().float()
                wv_hi = wv.half()
                v_w3 = f16(torch.cat([wv_hi, wv_hi, (wv - wv_hi.float()).half()], dim=1))
            else:
                qkv_w, qkv_b = m.in_proj_weight, m.in_proj_bias
                fc_w, fc_b, slope_p = f16(m.out_proj.weight), f32(m.out_proj.bias), None
                v_w3 = None
            layers[i] = StampMilLayer(f32(att.norm.weight), f32(att.norm.bias), f16(qkv_w), f32(qkv_b),
                                      v_w3, slope_p, fc_w, fc_b, f32(ff[0].weight), f32(ff[0].bias),
                                      f16(ff[1].weight), f32(ff[1].bias), f16(ff[4].weight), f32(ff[4].bias))
        self._packed = (key, keep, cfg, w, layers)
        return self._packed

    def _pack_padded(self, key, plan):
        """Same operands as :meth:`_pack` for a model outside the envelope: every weight zero-padded to the kernel
        shapes.  Residual-space vectors / matrix sides (LayerNorm, biases, class token, projection rows, fc rows,
        feed-forward) are padded at the end; head-space sides (q | k | v rows, fc columns) head by head, so that head
        ``h`` occupies columns ``[h * hp, h * hp + hd)``.  Zero rows / columns keep every padding channel exactly zero
        through GELU(0) = 0, the attention and the residual adds; LayerNorm statistics and the softmax scale use the
        real sizes (``dim_model_real``, ``head_dim_real``)."""
        import torch.nn.functional as Fn

        keep: list[Tensor] = []
        d, hd, hp, dc, F, Fp, ff, ffp = (plan[k] for k in ("d", "hd", "hp", "dc", "F", "Fp", "ff", "ffp"))
        H, use_alibi = self._cfg["n_heads"], bool(self._cfg["use_alibi"])

        def f32(t: Tensor) -> int:
            t = t.detach().float().contiguous()
            keep.append(t)
            return t.data_ptr()

        def f16(t: Tensor) -> int:
            t = t.detach().half().contiguous()
            keep.append(t)
            return t.data_ptr()

        vec = lambda v, n: Fn.pad(v.detach().float(), (0, n - v.shape[0]))               # residual / hidden vectors
        mat = lambda m, r, c: Fn.pad(m.detach().float(), (0, c - m.shape[1], 0, r - m.shape[0]))
        head_rows = lambda m: Fn.pad(m.detach().float().reshape(H, hd, -1), (0, 0, 0, hp - hd)).reshape(H * hp, -1)
        head_vec = lambda v: Fn.pad(v.detach().float().reshape(H, hd), (0, hp - hd)).reshape(H * hp)
        head_cols = lambda m: Fn.pad(m.detach().float().reshape(m.shape[0], H, hd), (0, hp - hd)).reshape(m.shape[0], H * hp)

        cfg = StampMilConfig(**{**self._cfg, "dim_input": Fp, "dim_model": dc, "dim_ff": ffp,
                                "dim_model_real": d, "head_dim_real": hd})
        w = StampMilWeights(f16(mat(self.project_features[0].weight, dc, Fp)), f32(vec(self.project_features[0].bias, dc)),
                            f32(vec(self.class_token, dc)), f32(vec(self.transformer.norm.weight, dc)),
                            f32(vec(self.transformer.norm.bias, dc)), f32(mat(self.mlp_head[0].weight, self._cfg["dim_output"], dc)),
                            f32(self.mlp_head[0].bias))
        layers = (StampMilLayer * max(1, self._cfg["n_layers"]))()
        for i, (att, ffn) in enumerate(self.transformer.layers):
            m = att.mhsa
            if use_alibi:
                groups = (m.query_encoders, m.key_encoders, m.value_encoders)
                qkv_w = torch.cat([mat(head_rows(torch.cat([e.weight for e in grp])), H * hp, dc) for grp in groups])
                qkv_b = torch.cat([head_vec(torch.cat([e.bias for e in grp])) for grp in groups])
                slope = torch.cat([a.bias_scale / a.scale_distance.running_mean for a in m.attentions])
                fc_full = mat(head_cols(m.fc.weight), dc, dc)
                fc_hi = round_to_tf32(fc_full)
                fc_lo = round_to_tf32(fc_full - fc_hi)
                fc_w, fc_b, slope_p = f32(torch.cat([fc_hi, fc_hi, fc_lo], dim=1)), f32(vec(m.fc.bias, dc)), f32(slope)
                wv = qkv_w[2 * dc:]
                wv_hi = wv.half()
                v_w3 = f16(torch.cat([wv_hi, wv_hi, (wv - wv_hi.float()).half()], dim=1))
            else:
                wq, wk, wv = m.in_proj_weight.detach().float().split(d)
                bq, bk, bv = m.in_proj_bias.detach().float().split(d)
                qkv_w = torch.cat([mat(head_rows(x), H * hp, dc) for x in (wq, wk, wv)])
                qkv_b = torch.cat([head_vec(x) for x in (bq, bk, bv)])
                fc_w, fc_b, slope_p = f16(mat(head_cols(m.out_proj.weight), dc, dc)), f32(vec(m.out_proj.bias, dc)), None
                v_w3 = None
            layers[i] = StampMilLayer(f32(vec(att.norm.weight, dc)), f32(vec(att.norm.bias, dc)), f16(qkv_w), f32(qkv_b),
                                      v_w3, slope_p, fc_w, fc_b, f32(vec(ffn[0].weight, dc)), f32(vec(ffn[0].bias, dc)),
                                      f16(mat(ffn[1].weight, ffp, dc)), f32(vec(ffn[1].bias, ffp)),
                                      f16(mat(ffn[4].weight, dc, ffp)), f32(vec(ffn[4].bias, dc)))
        return (key, keep, cfg, w, layers)

    def supports_ragged(self) -> bool:
        """True when bags of different lengths can go through ONE forward (``forward_ragged``): the unmasked inference
        path with head dimension 64 (native or zero-padded), which is the reference's default configuration."""
        return self._shape_plan()["hp"] == 64

    def pack_ragged(self, bags: list[tuple[Tensor, Tensor]], pin: bool = True) -> tuple[Tensor, Tensor, Tensor, int]:
        """[(feats [N_b, F], coords [N_b, 2])] on the host -> the operands of ``forward_ragged``: fp16 tokens
        [total, F'] with a placeholder row in front of every bag (the class-token row), fp32 coordinates [total, 2]
        with (0, 0) there, int32 row offsets [B + 1], and the longest bag in tokens.  ``F'`` is the kernels' input
        width (zero-padded when the model's is not a multiple of 8)."""
        _, _, cfg, _, _ = self._pack()
        sizes = [int(f.shape[0]) + 1 for f, _ in bags]
        off = [0]
        for n in sizes:
            off.append(off[-1] + n)
        total, Fin = off[-1], self._cfg["dim_input"]
        tokens = torch.empty((total, cfg.dim_input), dtype=torch.float16, pin_memory=pin)
        coords = torch.empty((total, 2), dtype=torch.float32, pin_memory=pin)
        if cfg.dim_input != Fin:
            tokens[:, Fin:] = 0
        for (f, c), o, n in zip(bags, off, sizes):
            if f.dim() != 2 or f.shape[1] != Fin or tuple(c.shape) != (n - 1, 2):
                raise TypeError(f"expected feats [N,{Fin}] and coords [N,2], got {tuple(f.shape)} / {tuple(c.shape)}")
            tokens[o] = 0                      # placeholder row of the class token (any finite values)
            coords[o] = 0
            tokens[o + 1:o + n, :Fin] = f
            coords[o + 1:o + n] = c
        return tokens, coords, torch.tensor(off, dtype=torch.int32), max(sizes)

    @torch.no_grad()
    def forward_ragged(self, tokens: Tensor, coords: Tensor, seq_off: Tensor, s_max: int) -> Tensor:
        """ONE inference forward over a ragged batch of bags (``pack_ragged`` layout, on the model's device): the dense
        layers run over the rows of all bags at once, the attention kernel walks each bag separately ->
        logits fp32 [B, dim_output], row b identical to ``forward(bag_b[None], coords=..., mask=None)``.
        What ``_predict`` (src/stamp/modeling/deploy.py:390-456) does patient by patient."""
        if not tokens.is_cuda or not self.class_token.is_cuda:
            raise RuntimeError("stamp_b200 VisionTransformer runs on a CUDA device only (no CPU fallback)")
        if not self.supports_ragged():
            raise ValueError("ragged batches need head dimension 64 (use forward per bag)")
        lib = _bind()
        _, _, cfg, w, layers = self._pack()
        if tokens.dtype != torch.float16 or tokens.dim() != 2 or tokens.shape[1] != cfg.dim_input or not tokens.is_contiguous():
            raise TypeError(f"tokens must be a contiguous fp16 [rows, {cfg.dim_input}] tensor (see pack_ragged)")
        total, B = tokens.shape[0], seq_off.numel() - 1
        if coords.dtype != torch.float32 or tuple(coords.shape) != (total, 2) or not coords.is_contiguous() or \
                seq_off.dtype != torch.int32 or not seq_off.is_cuda or B < 1:
            raise TypeError("coords must be fp32 [rows, 2], seq_off an int32 CUDA tensor [B + 1]")
        dev = tokens.device
        logits = torch.empty((B, self._cfg["dim_output"]), dtype=torch.float32, device=dev)
        need = lib.stamp_mil_ragged_workspace_bytes(C.byref(cfg), B, total, int(s_max))
        if need == 0:
            raise ValueError("unsupported MIL configuration / batch for the ragged forward")
        stream = torch.cuda.current_stream(dev).cuda_stream
        if self._workspace is None or len(self._workspace) > 8:
            self._workspace = {}
        ws = self._workspace.get(stream)
        if ws is None or ws.numel() < need or ws.device != dev:
            ws = self._workspace[stream] = torch.empty(need, dtype=torch.uint8, device=dev)
        code = lib.stamp_mil_forward_ragged(C.byref(cfg), C.byref(w), layers, tokens.data_ptr(), coords.data_ptr(),
                                            seq_off.data_ptr(), B, total, int(s_max), logits.data_ptr(), ws.data_ptr(),
                                            ws.numel(), stream)
        _lib.check(code, "stamp_mil_forward_ragged")
        return logits

    def forward(self, bags: Tensor, *, coords: Tensor, mask: Tensor | None) -> Tensor:
        if bags.dim() != 3 or coords.dim() != 3 or coords.shape[:2] != bags.shape[:2] or coords.shape[2] != 2:
            raise TypeError(f"expected bags [B,N,F] and coords [B,N,2], got {tuple(bags.shape)} / {tuple(coords.shape)}")
        if not bags.is_floating_point() or not coords.is_floating_point():
            raise TypeError("bags and coords must be floating point tensors")
        if mask is not None and (mask.dtype != torch.bool or mask.shape != bags.shape[:2]):
            raise TypeError("mask must be a bool tensor [B,N] or None")
        if bags.shape[2] != self._cfg["dim_input"]:
            raise ValueError(f"bags have {bags.shape[2]} features, model expects {self._cfg['dim_input']}")
        if not bags.is_cuda or not self.class_token.is_cuda:
            raise RuntimeError("stamp_b200 VisionTransformer runs on a CUDA device only (no CPU fallback)")
        wants_grad = torch.is_grad_enabled() and (bags.requires_grad or any(p.requires_grad for p in self.parameters()))
        if wants_grad and mask is not None:
            raise NotImplementedError(
                "stamp_b200 VisionTransformer: a masked forward cannot record an autograd graph (gradients exist "
                "for the mask=None branch, the one every Lightning step and the heatmap Jacobian take). Wrap "
                "masked evaluation in torch.no_grad() / inference_mode(), or freeze the model's parameters.")
        if mask is None and (wants_grad or self.training):
            # training_step / heatmap gradients: checkpointing forward + backward kernels (train.py).  A
            # train-mode forward under no_grad takes the same path, so that dropout and the running-mean update
            # of training mode happen exactly when the reference's module would apply them.
            from .train import mil_forward_with_grad
            return mil_forward_with_grad(self, bags, coords)
        lib = _bind()
        _, _, cfg, w, layers = self._pack()
        B, N, _ = bags.shape
        dev = bags.device
        # fp16 features (the dtype the .h5 feature files hold) are consumed as they are; anything else as fp32
        bags_in = bags.detach().contiguous() if bags.dtype == torch.float16 else bags.detach().float().contiguous()
        coords32 = coords.detach().float().contiguous()
        mask8 = mask.to(torch.uint8).contiguous() if mask is not None else None
        logits = torch.empty((B, self._cfg["dim_output"]), dtype=torch.float32, device=dev)
        if cfg.dim_input != bags_in.shape[2]:       # zero-padded input width (see _shape_plan)
            bags_in = torch.nn.functional.pad(bags_in, (0, cfg.dim_input - bags_in.shape[2]))
        need = lib.stamp_mil_workspace_bytes(C.byref(cfg), B, N)
        if need == 0:
            raise ValueError("unsupported MIL configuration for the sm_100a kernels")
        stream = torch.cuda.current_stream(dev).cuda_stream
        if self._workspace is None or len(self._workspace) > 8:   # (callers that keep creating streams: start over)
            self._workspace = {}
        ws = self._workspace.get(stream)
        if ws is None or ws.numel() < need or ws.device != dev:
            ws = self._workspace[stream] = torch.empty(need, dtype=torch.uint8, device=dev)
        code = lib.stamp_mil_forward(C.byref(cfg), C.byref(w), layers, bags_in.data_ptr(),
                                     int(bags_in.dtype == torch.float16), coords32.data_ptr(),
                                     None if mask8 is None else mask8.data_ptr(), logits.data_ptr(), B, N,
                                     ws.data_ptr(), ws.numel(), stream)
        _lib.check(code, "stamp_mil_forward")
        # fp16 bags: logits stay in the model's fp32 (predict_step casts the bags to the parameter dtype instead)
        return logits if bags.dtype == torch.float16 else logits.to(bags.dtype)
