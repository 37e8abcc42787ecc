"""MIL training step on the B200 path: autograd binding of ``stamp_mil_train_forward/backward``,
the Lightning ``_step`` loss, the running-mean statistic of training mode and a fused AdamW.

Mirrors, in the reference (paths relative to its root):

* ``LitTileClassifier._step`` / ``training_step``  src/stamp/modeling/models/__init__.py:239-286
  -> :func:`training_step`, :func:`cross_entropy`
* ``Base.configure_optimizers``                    src/stamp/modeling/models/__init__.py:133-141
  -> :func:`configure_optimizers` (``FusedAdamW`` + ``torch.optim.lr_scheduler.OneCycleLR``)
* ``_RunningMeanScaler.forward`` (training branch)  src/stamp/modeling/models/vision_tranformer.py:23-31
  -> :func:`update_running_means`
* ``_gradcam_per_category``                         src/stamp/heatmaps/__init__.py:36-56
  -> :func:`gradcam_per_category`

All arithmetic runs in ``libstamp_b200.so``; torch is used for memory, streams and autograd graph
plumbing (``torch.cat`` of the per-head Linear weights routes the packed gradients back to the
reference's individual parameters).  There is no CPU fallback.
"""

from __future__ import annotations

import ctypes as C
from typing import Iterable

import torch
from torch import Tensor

from . import _lib
from .mil import StampMilConfig, VisionTransformer, bump_weights_epoch

_TOP_FIELDS = ("proj_w", "proj_b", "class_token", "norm_w", "norm_b", "head_w", "head_b")
_LAYER_FIELDS = ("ln1_w", "ln1_b", "qkv_w", "qkv_b", "bias_scale", "fc_w", "fc_b", "ln2_w", "ln2_b",
                 "ff1_w", "ff1_b", "ff2_w", "ff2_b")


class StampMilTrainTop(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in _TOP_FIELDS]


class StampMilTrainLayer(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in _LAYER_FIELDS]


class StampMilTrainStep(C.Structure):
    _fields_ = [("p_drop_proj", C.c_float), ("p_drop_ff", C.c_float), ("seed", C.c_ulonglong),
                ("inv_rm", C.c_void_p)]


def _bind() -> C.CDLL:
    lib = _lib.load()
    if getattr(lib, "_train_bound", False):
        return lib
    P = C.POINTER
    vp, i32, f32, sz, ll = C.c_void_p, C.c_int, C.c_float, C.c_size_t, C.c_longlong
    lib.stamp_mil_train_ctx_bytes.restype = sz
    lib.stamp_mil_train_ctx_bytes.argtypes = [P(StampMilConfig), i32, i32]
    lib.stamp_mil_train_forward.restype = i32
    lib.stamp_mil_train_forward.argtypes = [P(StampMilConfig), P(StampMilTrainTop), P(StampMilTrainLayer),
                                            P(StampMilTrainStep), vp, vp, vp, i32, i32, vp, sz, vp]
    lib.stamp_mil_train_backward.restype = i32
    lib.stamp_mil_train_backward.argtypes = [P(StampMilConfig), P(StampMilTrainTop), P(StampMilTrainLayer),
                                             P(StampMilTrainStep), vp, P(StampMilTrainTop), P(StampMilTrainLayer),
                                             vp, i32, i32, vp, sz, vp]
    lib.stamp_mil_train_dropout_mask.restype = i32
    lib.stamp_mil_train_dropout_mask.argtypes = [C.c_ulonglong, i32, ll, f32, vp, vp]
    lib.stamp_pairwise_dist_mean_workspace_bytes.restype = sz
    lib.stamp_pairwise_dist_mean_workspace_bytes.argtypes = [i32, i32]
    lib.stamp_pairwise_dist_mean.restype = i32
    lib.stamp_pairwise_dist_mean.argtypes = [vp, i32, i32, vp, vp, sz, vp]
    lib.stamp_cross_entropy.restype = i32
    lib.stamp_cross_entropy.argtypes = [vp, vp, vp, i32, i32, f32, vp, vp, vp]
    lib.stamp_cox_loss.restype = i32
    lib.stamp_cox_loss.argtypes = [vp, vp, vp, i32, i32, f32, vp, vp, vp]
    lib.stamp_l1_loss.restype = i32
    lib.stamp_l1_loss.argtypes = [vp, vp, ll, f32, vp, vp, vp]
    lib.stamp_adamw_step.restype = i32
    lib.stamp_adamw_step.argtypes = [vp, vp, vp, vp, ll, f32, f32, f32, f32, f32, i32, f32, vp]
    lib._train_bound = True
    return lib


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(t: Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{what} must live on a CUDA device: the B200 training path has no CPU fallback")


# ---- training-mode running mean ------------------------------------------------------------------
def pairwise_dist_mean(coords: Tensor) -> Tensor:
    """Mean over [B, S, S] of the token distances (class token at (0,0) included), device scalar."""
    _need_cuda(coords, "coords")
    lib = _bind()
    B, N, _ = coords.shape
    c = coords.detach().float().contiguous()
    ws = torch.empty(lib.stamp_pairwise_dist_mean_workspace_bytes(B, N), dtype=torch.uint8, device=c.device)
    out = torch.empty(1, dtype=torch.float32, device=c.device)
    _lib.check(lib.stamp_pairwise_dist_mean(c.data_ptr(), B, N, out.data_ptr(), ws.data_ptr(), ws.numel(),
                                            _stream()), "stamp_pairwise_dist_mean")
    return out


def _scalers(model: VisionTransformer):
    return [a.scale_distance for att, _ in model.transformer.layers for a in att.mhsa.attentions]


@torch.no_grad()
def update_running_means(model: VisionTransformer, coords: Tensor) -> None:
    """rm <- mean(rm + (dist - rm) / n) = rm + (mean(dist) - rm) / n;  n += 1  for every head's scaler,
    without a host synchronisation."""
    sc = _scalers(model)
    if not sc:
        return
    m = pairwise_dist_mean(coords)
    rms = [s.running_mean for s in sc]
    ns = [s.items_so_far for s in sc]
    rm = torch.cat(rms)
    new = rm + (m - rm) / torch.cat(ns)
    torch._foreach_copy_(rms, list(new.split(1)))
    torch._foreach_add_(ns, 1.0)


# ---- autograd binding ------------------------------------------------------------------------------
def _packed_params(model: VisionTransformer) -> list[Tensor]:
    """fp32 tensors in C-struct order (7 + 13 per layer).  The per-head Linear weights are concatenated
    with differentiable torch.cat, so autograd splits the packed gradients back onto the reference's
    parameters."""
    out = [model.project_features[0].weight, model.project_features[0].bias, model.class_token,
           model.transformer.norm.weight, model.transformer.norm.bias, model.mlp_head[0].weight,
           model.mlp_head[0].bias]
    H = model._cfg["n_heads"]
    for att, ff in model.transformer.layers:
        m = att.mhsa
        if model._cfg["use_alibi"]:
            groups = (m.query_encoders, m.key_encoders, m.value_encoders)
            attn = [torch.cat([e.weight for grp in groups for e in grp]),
                    torch.cat([e.bias for grp in groups for e in grp]),
                    torch.cat([a.bias_scale for a in m.attentions]), m.fc.weight, m.fc.bias]
        else:   # nn.MultiheadAttention: in_proj rows are already q | k | v with the heads side by side
            attn = [m.in_proj_weight, m.in_proj_bias, torch.zeros(H, device=m.in_proj_weight.device),
                    m.out_proj.weight, m.out_proj.bias]
        out += [att.norm.weight, att.norm.bias, *attn, ff[0].weight, ff[0].bias, ff[1].weight, ff[1].bias,
                ff[4].weight, ff[4].bias]
    return out


def _packed_groups(model: VisionTransformer) -> list[list[Tensor] | None]:
    """The same 7 + 13 per layer slots as :func:`_packed_params`, as groups of parameters whose row-wise concatenation
    IS the packed tensor (None: a slot without parameters -- the bias_scale placeholder of nn.MultiheadAttention).
    An optimiser that lays its flat buffers out in this order (``FusedAdamW(..., model=model)``) makes every packed
    tensor a contiguous view: no torch.cat in the forward, and the backward kernels accumulate straight into
    ``flat_grad``."""
    out: list[list[Tensor] | None] = [[model.project_features[0].weight], [model.project_features[0].bias],
                                      [model.class_token], [model.transformer.norm.weight], [model.transformer.norm.bias],
                                      [model.mlp_head[0].weight], [model.mlp_head[0].bias]]
    for att, ff in model.transformer.layers:
        m = att.mhsa
        if model._cfg["use_alibi"]:
            groups = (m.query_encoders, m.key_encoders, m.value_encoders)
            attn = [[e.weight for grp in groups for e in grp], [e.bias for grp in groups for e in grp],
                    [a.bias_scale for a in m.attentions], [m.fc.weight], [m.fc.bias]]
        else:
            attn = [[m.in_proj_weight], [m.in_proj_bias], None, [m.out_proj.weight], [m.out_proj.bias]]
        out += [[att.norm.weight], [att.norm.bias], *attn, [ff[0].weight], [ff[0].bias], [ff[1].weight], [ff[1].bias],
                [ff[4].weight], [ff[4].bias]]
    return out


def _structs(tensors: list[Tensor], n_layers: int):
    top = StampMilTrainTop(*[t.data_ptr() for t in tensors[:7]])
    layers = (StampMilTrainLayer * n_layers)()
    for l in range(n_layers):
        layers[l] = StampMilTrainLayer(*[t.data_ptr() for t in tensors[7 + 13 * l: 20 + 13 * l]])
    return top, layers


class _TrainState:
    """What one checkpointing forward leaves for its backward(s) (not a pytree: functorch passes it through)."""

    __slots__ = ("model", "cfg", "gen", "buf", "tensors", "inv_rm", "step_args", "shape", "bags_dtype", "sink")


def _run_backward(st: _TrainState, dlogits: Tensor, want_dbags: bool) -> tuple[Tensor | None, list[Tensor]]:
    lib = _bind()
    model, cfg = st.model, st.cfg
    if model._train_gen != st.gen:
        raise RuntimeError("the checkpoint buffer of this forward was overwritten by a later training-mode "
                           "forward of the same model; run backward before the next forward")
    B, N = st.shape
    if st.sink is not None and not want_dbags:
        # gradients accumulate straight into the optimiser's flat buffer (packed-order views of it)
        grads = st.sink
    else:
        sizes = [t.numel() for t in st.tensors]
        flat = torch.zeros(sum(sizes), dtype=torch.float32, device=dlogits.device)
        grads = [g.view_as(t) for g, t in zip(flat.split(sizes), st.tensors)]
    top, layers = _structs(st.tensors, cfg.n_layers)
    gtop, glayers = _structs(grads, cfg.n_layers)
    step = StampMilTrainStep(*st.step_args, st.inv_rm.data_ptr())
    dbags = torch.empty((B, N, cfg.dim_input), dtype=torch.float32, device=dlogits.device) if want_dbags else None
    dl = dlogits.detach().float().contiguous()
    _lib.check(lib.stamp_mil_train_backward(C.byref(cfg), C.byref(top), layers, C.byref(step), dl.data_ptr(),
                                            C.byref(gtop), glayers, None if dbags is None else dbags.data_ptr(),
                                            B, N, st.buf.data_ptr(), st.buf.numel(), _stream()),
               "stamp_mil_train_backward")
    return dbags, grads


class _MilBwdFn(torch.autograd.Function):
    """The backward as a Function of its own, so that functorch can vmap it: ``torch.func.jacrev(model)``
    (heatmaps, src/stamp/heatmaps/__init__.py:41-52) maps the C one-hot cotangents through :meth:`vmap`,
    which runs the C backward passes of the one checkpointed forward one after the other."""

    @staticmethod
    def forward(dlogits: Tensor, st: _TrainState, want_dbags: bool):
        dbags, grads = _run_backward(st, dlogits, want_dbags)
        if dbags is None:
            dbags = dlogits.new_zeros(0)
        if st.sink is not None and not want_dbags:
            return (dbags.to(st.bags_dtype),)        # the parameter gradients are already where they belong
        return (dbags.to(st.bags_dtype), *grads)

    @staticmethod
    def setup_context(ctx, inputs, output):
        pass

    @staticmethod
    def backward(ctx, *grads):
        raise NotImplementedError("double backward through the B200 MIL kernels is not implemented")

    @staticmethod
    def vmap(info, in_dims, dlogits, st, want_dbags):
        rows = [_MilBwdFn.apply(dlogits.select(in_dims[0], i), st, want_dbags) for i in range(info.batch_size)]
        return tuple(torch.stack(r) for r in zip(*rows)), tuple(0 for _ in rows[0])


class _MilTrainFn(torch.autograd.Function):
    @staticmethod
    def forward(model: VisionTransformer, bags: Tensor, coords: Tensor, inv_rm: Tensor, p_proj: float,
                p_ff: float, seed: int, *params: Tensor) -> Tensor:
        lib = _bind()
        cfg = StampMilConfig(**{**model._cfg, "dim_input": bags.shape[2]})    # (possibly zero-padded, see below)
        B, N, _ = bags.shape
        dev = bags.device
        need = lib.stamp_mil_train_ctx_bytes(C.byref(cfg), B, N)
        if need == 0:
            raise ValueError("unsupported MIL configuration for the sm_100a training kernels (needs head dim "
                             "32/64, dim_model and dim_feedforward % 8 == 0, dim_model <= 1024, at least one tile; "
                             "inference zero-pads other shapes, training only the input width)")
        buf = model._train_ctx
        if buf is None or buf.numel() < need or buf.device != dev:
            buf = model._train_ctx = torch.empty(need, dtype=torch.uint8, device=dev)
        model._train_gen += 1
        tensors = [p.detach().float().contiguous() for p in params]     # (views of flat_param: no copies)
        top, layers = _structs(tensors, cfg.n_layers)
        bags32 = bags.detach().float().contiguous()
        coords32 = coords.detach().float().contiguous()
        step = StampMilTrainStep(p_proj, p_ff, seed, inv_rm.data_ptr())
        logits = torch.empty((B, cfg.dim_output), dtype=torch.float32, device=dev)
        _lib.check(lib.stamp_mil_train_forward(C.byref(cfg), C.byref(top), layers, C.byref(step), bags32.data_ptr(),
                                               coords32.data_ptr(), logits.data_ptr(), B, N, buf.data_ptr(),
                                               buf.numel(), _stream()), "stamp_mil_train_forward")
        st = _TrainState()
        st.model, st.cfg, st.gen, st.buf = model, cfg, model._train_gen, buf
        st.tensors, st.inv_rm, st.step_args = tensors, inv_rm, (p_proj, p_ff, seed)
        st.shape, st.bags_dtype = (B, N), bags.dtype
        st.sink = model._grad_sink_views
        model._grad_sink_views = None
        model._train_state = st          # handed to setup_context (the functorch-compatible Function protocol)
        return logits

    @staticmethod
    def setup_context(ctx, inputs, output):
        # functorch calls this twice for one forward (once per transform level): the hand-over slot stays set
        ctx.st = inputs[0]._train_state

    @staticmethod
    def backward(ctx, dlogits: Tensor):
        want = bool(ctx.needs_input_grad[1])
        ctx.st.model._train_state = None    # the hand-over slot only; ctx keeps the state (no model<->state cycle)
        dbags, *grads = _MilBwdFn.apply(dlogits, ctx.st, want)
        if not grads:                       # gradient sink: nothing for autograd to route (see mil_forward_with_grad)
            grads = [None] * (len(ctx.needs_input_grad) - 7)
        return (None, dbags if want else None, None, None, None, None, None, *grads)


def mil_forward_with_grad(model: VisionTransformer, bags: Tensor, coords: Tensor) -> Tensor:
    """``model(bags, coords=coords, mask=None)`` with autograd: training mode applies the reference's
    dropouts and running-mean update, eval mode (heatmaps) neither."""
    alibi = bool(model._cfg["use_alibi"])
    if not alibi and model.training and any(att.mhsa.dropout > 0 for att, _ in model.transformer.layers):
        raise NotImplementedError("dropout on the attention weights of nn.MultiheadAttention (use_alibi=False with "
                                  "dropout > 0) is not implemented in the B200 training kernels")
    _need_cuda(bags, "bags")
    _need_cuda(model.class_token, "the model")
    if bags.shape[1] == 0:
        raise ValueError("empty bags cannot be trained on")
    L, H = model._cfg["n_layers"], model._cfg["n_heads"]
    if model.training:
        if alibi:
            update_running_means(model, coords)
        p_proj = float(model.project_features[2].p)
        p_ff = float(model.transformer.layers[0][1][3].p) if L > 0 else 0.0
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if (p_proj > 0 or p_ff > 0) else 0
    else:
        p_proj = p_ff = 0.0
        seed = 0
    with torch.no_grad():
        if alibi:
            inv_rm = (1.0 / torch.cat([s.running_mean for s in _scalers(model)]).float()).reshape(L, H).contiguous()
        else:
            inv_rm = torch.ones(L, H, device=bags.device)
    Fin = bags.shape[2]
    opt = model._flat_opt() if model._flat_opt is not None else None
    if (opt is not None and model.training and not bags.requires_grad and Fin % 8 == 0 and opt.owns(model)
            and not torch._C._functorch.is_functorch_wrapped_tensor(bags)):
        # parameters and gradients live in packed order inside the optimiser's flat buffers: the kernels read the
        # packed views directly (no torch.cat) and accumulate into flat_grad (no per-parameter adds by autograd).
        # class_token rides along as the differentiable input that makes autograd call the backward.
        packed, model._grad_sink_views = opt.packed_views()
        return _MilTrainFn.apply(model, bags, coords, inv_rm, p_proj, p_ff, seed, *packed, model.class_token).to(bags.dtype)
    params = _packed_params(model)
    if Fin % 8:
        # input widths that are not a multiple of 8 (the reference's acceptance test trains on 25 features): zero
        # columns on both sides of the projection; autograd slices the padded gradients back
        pad = (Fin + 7) // 8 * 8 - Fin
        bags_p = torch.nn.functional.pad(bags, (0, pad))
        params[0] = torch.nn.functional.pad(params[0], (0, pad))
        return _MilTrainFn.apply(model, bags_p, coords, inv_rm, p_proj, p_ff, seed, *params).to(bags.dtype)
    return _MilTrainFn.apply(model, bags, coords, inv_rm, p_proj, p_ff, seed, *params).to(bags.dtype)


def dropout_keep_mask(seed: int, site: int, n: int, p: float, device) -> Tensor:
    """The keep mask the kernels use at a dropout site (tests build the oracle's masks from it)."""
    lib = _bind()
    out = torch.empty(n, dtype=torch.uint8, device=device)
    _lib.check(lib.stamp_mil_train_dropout_mask(seed, site, n, p, out.data_ptr(), _stream()),
               "stamp_mil_train_dropout_mask")
    return out


# ---- loss --------------------------------------------------------------------------------------------
class _CrossEntropyFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits: Tensor, targets: Tensor, class_weights: Tensor | None) -> Tensor:
        lib = _bind()
        _need_cuda(logits, "logits")
        B, Cn = logits.shape
        l32 = logits.detach().float().contiguous()
        t32 = targets.detach().to(l32.device).float().contiguous()
        w32 = None if class_weights is None else class_weights.detach().to(l32.device).float().contiguous()
        loss = torch.empty((), dtype=torch.float32, device=l32.device)
        dl = torch.empty_like(l32)
        _lib.check(lib.stamp_cross_entropy(l32.data_ptr(), t32.data_ptr(), None if w32 is None else w32.data_ptr(),
                                           B, Cn, 1.0, loss.data_ptr(), dl.data_ptr(), _stream()),
                   "stamp_cross_entropy")
        ctx.save_for_backward(dl)
        return loss

    @staticmethod
    def backward(ctx, g: Tensor):
        (dl,) = ctx.saved_tensors
        return dl * g, None, None


def cross_entropy(logits: Tensor, targets: Tensor, class_weights: Tensor | None = None) -> Tensor:
    """``F.cross_entropy(logits, soft_targets, weight=class_weights)`` as the reference's ``_step`` calls it."""
    return _CrossEntropyFn.apply(logits, targets, class_weights)


def training_step(model: VisionTransformer, batch, class_weights: Tensor | None = None) -> Tensor:
    """``LitTileClassifier._step(step_name='training', use_mask=False)``: batch = (bags, coords, bag_sizes, targets)."""
    bags, coords, _bag_sizes, targets = batch
    logits = model(bags, coords=coords, mask=None)
    return cross_entropy(logits, targets.to(logits.dtype), class_weights)


class _L1LossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred: Tensor, target: Tensor) -> Tensor:
        lib = _bind()
        _need_cuda(pred, "pred")
        p32 = pred.detach().float().contiguous()
        t32 = target.detach().to(p32.device).float().expand_as(p32).contiguous()
        loss = torch.empty((), dtype=torch.float32, device=p32.device)
        dp = torch.empty_like(p32)
        _lib.check(lib.stamp_l1_loss(p32.data_ptr(), t32.data_ptr(), p32.numel(), 1.0, loss.data_ptr(), dp.data_ptr(), _stream()),
                   "stamp_l1_loss")
        ctx.save_for_backward(dp)
        return loss

    @staticmethod
    def backward(ctx, g: Tensor):
        (dp,) = ctx.saved_tensors
        return dp * g, None


def l1_loss(pred: Tensor, target: Tensor) -> Tensor:
    """``nn.functional.l1_loss`` as ``LitBaseRegressor._compute_loss`` calls it (models/__init__.py:420-422); the gradient
    goes to ``pred`` (the reference passes its arguments in the order (preds, y), the loss is symmetric)."""
    return _L1LossFn.apply(pred, target)


class _CoxLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, log_hz: Tensor, time: Tensor, event: Tensor, breslow: bool) -> Tensor:
        lib = _bind()
        _need_cuda(log_hz, "log_hz")
        s32 = log_hz.detach().float().contiguous()
        t32 = time.detach().to(s32.device).float().contiguous()
        e8 = (event.detach().to(s32.device) != 0).to(torch.uint8).contiguous()
        if not (s32.ndim == 1 and t32.shape == s32.shape and e8.shape == s32.shape):
            raise ValueError(f"log_hz, time and event must be vectors of one length, got {tuple(log_hz.shape)}, "
                             f"{tuple(time.shape)}, {tuple(event.shape)}")
        loss = torch.empty((), dtype=torch.float32, device=s32.device)
        ds = torch.empty_like(s32)
        _lib.check(lib.stamp_cox_loss(s32.data_ptr(), t32.data_ptr(), e8.data_ptr(), s32.numel(), int(breslow), 1.0, loss.data_ptr(),
                                      ds.data_ptr(), _stream()), "stamp_cox_loss")
        ctx.save_for_backward(ds)
        return loss

    @staticmethod
    def backward(ctx, g: Tensor):
        (ds,) = ctx.saved_tensors
        return ds * g, None, None, None


def neg_partial_log_likelihood(log_hz: Tensor, time: Tensor, event: Tensor, ties_method: str = "efron",
                               reduction: str = "mean") -> Tensor:
    """``neg_partial_log_likelihood`` of the reference (models/cox.py:107-268) in one launch, no host synchronisation:
    Cox partial likelihood, Efron's (default) or Breslow's handling of tied times, reduction "mean".  Where the reference
    returns a fresh zero leaf (no events in the batch) this returns a zero that is still connected to ``log_hz`` (zero
    gradient); NaN inputs are not filtered (the reference's ``nanmean``)."""
    if ties_method not in ("efron", "breslow"):
        raise ValueError(f'Ties method {ties_method} should be one of ["efron", "breslow"]')
    if reduction.lower() != "mean":
        raise ValueError(f"Reduction {reduction} is not implemented, only 'mean' (what the reference's steps use)")
    if log_hz.ndim == 0:
        return log_hz * 0.0
    return _CoxLossFn.apply(log_hz, time, event, ties_method == "breslow")


def regression_step(model: VisionTransformer, batch) -> Tensor:
    """``LitTileRegressor._step(step_name='training', use_mask=False)`` (models/__init__.py:444-462): dim_output = 1,
    batch = (bags, coords, bag_sizes, targets [B, 1])."""
    bags, coords, _bag_sizes, targets = batch
    preds = model(bags, coords=coords, mask=None)
    return l1_loss(preds, targets.to(preds.device).float())


def survival_step(model: VisionTransformer, batch) -> Tensor:
    """``LitTileSurvival.training_step`` (models/__init__.py:751-776): dim_output = 1, targets [B, 2] = (time, event)."""
    bags, coords, _bag_sizes, targets = batch
    preds = model(bags, coords=coords, mask=None)
    y = targets.to(preds.device, dtype=torch.float32)
    return neg_partial_log_likelihood(preds.squeeze(-1), y[:, 0], y[:, 1])


# ---- optimiser -----------------------------------------------------------------------------------------
class FusedAdamW(torch.optim.Optimizer):
    """``torch.optim.AdamW`` semantics in one kernel launch per step over a flat fp32 parameter buffer.

    The parameters are re-pointed into one contiguous buffer (``p.data`` become views) and so are their
    ``.grad``; ``flat_grad`` is also what :class:`stamp_b200.sharding.FlatGradAllReducer` all-reduces.
    Being a ``torch.optim.Optimizer`` it works with ``OneCycleLR`` (which cycles ``lr`` and ``betas[0]``).
    One parameter group only; the moments and the step count travel in ``state_dict()`` (key ``"fused"``).

    Built with ``model=``, the training backward of that model adds its parameter gradients into ``flat_grad`` itself
    instead of handing them to autograd (as long as every ``p.grad`` is still the view this class installed):
    ``loss.backward()`` then behaves as usual, gradient accumulation over several backwards included, but
    ``torch.autograd.grad(loss, parameters)`` and per-parameter autograd hooks see nothing -- build the optimiser
    without ``model=`` where those are needed."""

    def __init__(self, params: Iterable[Tensor], lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-2, model: VisionTransformer | None = None) -> None:
        """``model``: lay the flat buffers out in the order the training kernels consume the parameters (per-head
        Linear weights of a layer adjacent, :func:`_packed_groups`), so that the packed operands are views of
        ``flat_param`` and the backward accumulates straight into ``flat_grad``.  ``params`` must then be exactly the
        model's trainable parameters."""
        params = list(params)
        if any(isinstance(p, dict) for p in params):
            raise ValueError("FusedAdamW keeps ONE flat buffer: parameter groups are not supported")
        params = [p for p in params if p.requires_grad]
        if not params:
            raise ValueError("FusedAdamW got no trainable parameters")
        for p in params:
            _need_cuda(p, "parameters")
            if p.dtype != torch.float32:
                raise TypeError("FusedAdamW keeps fp32 master parameters")
        self._built = False
        self._groups: list[tuple[int, tuple[int, ...]] | None] | None = None
        if model is not None:
            groups = _packed_groups(model)
            ordered = [p for g in groups if g is not None for p in g]
            if len(ordered) != len(params) or {id(p) for p in ordered} != {id(p) for p in params}:
                raise ValueError("FusedAdamW(model=...) needs exactly the model's trainable parameters")
            params = ordered
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        dev = params[0].device
        # every parameter (with a model: every packed group) starts on a 256-byte boundary: the GEMM / vector kernels
        # need 16-byte aligned bases; inside a group the members follow each other without gaps
        sizes = [p.numel() for p in params]
        offs, total = [], 0
        if model is None:
            for n in sizes:
                offs.append(total)
                total += (n + 63) // 64 * 64
        else:
            self._groups = []
            for g in groups:
                if g is None:
                    self._groups.append(None)
                    continue
                rows = sum(p.shape[0] if p.dim() > 0 else 1 for p in g)
                self._groups.append((total, (rows, *g[0].shape[1:])))
                for p in g:
                    offs.append(total)
                    total += p.numel()
                total = (total + 63) // 64 * 64
        self.flat_param = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros_like(self.flat_param)
        self.exp_avg = torch.zeros_like(self.flat_param)
        self.exp_avg_sq = torch.zeros_like(self.flat_param)
        self._params, self._offs, self._sizes = params, offs, sizes
        with torch.no_grad():
            for p, o, n in zip(params, offs, sizes):
                view = self.flat_param[o:o + n]
                view.copy_(p.data.reshape(-1))
                p.data = view.view_as(p)
                p.grad = self.flat_grad[o:o + n].view_as(p)
        self._step = 0
        self._built = True
        if model is not None:
            import weakref

            self._model_ref = weakref.ref(model)
            self._dummy = torch.zeros(64, dtype=torch.float32, device=dev)   # slot without parameters (MHA bias_scale)
            model._flat_opt = weakref.ref(self)
        bump_weights_epoch()

    def owns(self, model) -> bool:
        """True while ``model``'s parameters are still the views of ``flat_param`` this optimiser created."""
        if self._groups is None or getattr(self, "_model_ref", lambda: None)() is not model:
            return False
        pbase, gbase = self.flat_param.data_ptr(), self.flat_grad.data_ptr()
        for p, o in zip(self._params, self._offs):
            g = p.grad
            if g is None or p.data_ptr() != pbase + 4 * o or g.data_ptr() != gbase + 4 * o:
                return False
        return True

    def packed_views(self) -> tuple[list[Tensor], list[Tensor]]:
        """(packed parameter operands, matching gradient accumulators) as views of the flat buffers, in the order of
        the training kernels' structs."""
        ps, gs = [], []
        for g in self._groups:
            if g is None:
                ps.append(self._dummy)
                gs.append(self._dummy)
                continue
            off, shape = g
            n = 1
            for d in shape:
                n *= d
            ps.append(self.flat_param[off:off + n].view(shape))
            gs.append(self.flat_grad[off:off + n].view(shape))
        return ps, gs

    def add_param_group(self, param_group) -> None:
        if getattr(self, "_built", False):
            raise ValueError("FusedAdamW keeps ONE flat buffer: parameter groups cannot be added after construction")
        super().add_param_group(param_group)

    # ---- checkpointing: the moments live outside Optimizer.state ---------------------------------------
    def state_dict(self) -> dict:
        sd = super().state_dict()
        sd["fused"] = {"step": self._step, "exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone(),
                       "sizes": list(self._sizes)}
        return sd

    def load_state_dict(self, state_dict: dict) -> None:
        fused = state_dict.get("fused")
        if fused is None:
            raise KeyError("not a FusedAdamW state dict (no 'fused' entry): the Adam moments would silently restart")
        if list(fused["sizes"]) != list(self._sizes):
            raise ValueError("FusedAdamW state dict was saved for a different parameter list")
        super().load_state_dict({k: v for k, v in state_dict.items() if k != "fused"})
        with torch.no_grad():
            self.exp_avg.copy_(fused["exp_avg"])
            self.exp_avg_sq.copy_(fused["exp_avg_sq"])
        self._step = int(fused["step"])

    # ---- the parameters and their gradients must still be views of the flat buffers ---------------------
    @torch.no_grad()
    def relink(self) -> None:
        """``model.zero_grad()`` (``set_to_none=True``), ``p.grad = None`` or a stray ``.grad`` assignment detach a
        gradient from ``flat_grad``: fold whatever autograd accumulated elsewhere back in and re-link the view.
        A parameter that no longer lives in ``flat_param`` (``model.half()`` / ``.to()`` after construction) is an
        error: the kernel would update memory the model does not read."""
        pbase, gbase = self.flat_param.data_ptr(), self.flat_grad.data_ptr()
        for p, o, n in zip(self._params, self._offs, self._sizes):
            if p.data_ptr() != pbase + 4 * o or p.dtype != torch.float32:
                raise RuntimeError("a parameter was moved or cast after FusedAdamW was built (its storage is no longer "
                                   "the optimiser's flat buffer); build the optimiser after model.to(...)")
            g = p.grad
            if g is None or g.data_ptr() != gbase + 4 * o:
                view = self.flat_grad[o:o + n].view_as(p)
                if g is None:
                    view.zero_()
                else:
                    view.copy_(g)
                p.grad = view

    def zero_grad(self, set_to_none: bool = False) -> None:  # grads stay views of flat_grad
        self.flat_grad.zero_()
        self.relink()

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        self.relink()
        lib = _bind()
        g = self.param_groups[0]
        self._step += 1
        _lib.check(lib.stamp_adamw_step(self.flat_param.data_ptr(), self.flat_grad.data_ptr(),
                                        self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), self.flat_param.numel(),
                                        float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]),
                                        float(g["weight_decay"]), self._step, float(grad_scale), _stream()),
                   "stamp_adamw_step")
        # the kernel wrote through raw pointers, which no version counter sees: invalidate the packed-weight
        # caches (VisionTransformer._pack) explicitly
        bump_weights_epoch()
        return loss


def configure_optimizers(model: VisionTransformer, *, total_steps: int, max_lr: float = 1e-4,
                         div_factor: float = 25.0):
    """``Base.configure_optimizers`` (models/__init__.py:133-141): AdamW(lr=1e-3 placeholder) + OneCycleLR."""
    # all parameters trainable (the reference never freezes any): packed-order layout, gradients land in flat_grad
    packed = all(p.requires_grad for p in model.parameters())
    opt = FusedAdamW(model.parameters(), lr=1e-3, model=model if packed else None)
    sched = torch.optim.lr_scheduler.OneCycleLR(optimizer=opt, total_steps=total_steps, max_lr=max_lr,
                                                div_factor=div_factor)
    return opt, sched


def data_parallel_step(model: VisionTransformer, opt: FusedAdamW, batch, class_weights: Tensor | None = None,
                       scheduler=None) -> Tensor:
    """One optimiser step of data-parallel MIL training: every rank runs
    :func:`training_step` on its own bags, then ONE all-reduce of the flat gradient buffer (NCCL over
    NVLink; 14.7 MB for the default model) and the average of the ALiBi running means keep the replicas
    identical.  With a single process it is the plain training step.  Returns the local loss."""
    from .sharding import all_reduce_flat_sum, sync_alibi_running_mean

    opt.zero_grad()
    loss = training_step(model, batch, class_weights)
    loss.backward()
    opt.relink()                 # gradients that autograd put elsewhere are folded into flat_grad first
    scale = all_reduce_flat_sum(opt.flat_grad)
    sync_alibi_running_mean(model)
    opt.step(grad_scale=scale)
    if scheduler is not None:
        scheduler.step()
    return loss.detach()


# ---- heatmaps ---------------------------------------------------------------------------------------------
def gradcam_per_category(model: VisionTransformer, feats: Tensor, coords: Tensor) -> Tensor:
    """``_gradcam_per_category`` (src/stamp/heatmaps/__init__.py:36-56): cam[n, c] =
    softmax_n(|mean_d(feats * d logit_c / d feats)|).  The reference takes ``jacrev`` of the model; here
    the C rows of the Jacobian are C backward passes of one checkpointed forward."""
    was_training = model.training
    model.eval()
    try:
        x = feats.detach().float().requires_grad_(True)
        with torch.enable_grad():
            logits = mil_forward_with_grad(model, x[None], coords[None].float())[0]
            rows = []
            for c in range(logits.shape[0]):
                (j,) = torch.autograd.grad(logits[c], x, retain_graph=c + 1 < logits.shape[0])
                rows.append((x.detach() * j).mean(dim=-1).abs())
        return torch.softmax(torch.stack(rows), dim=-1).T
    finally:
        model.train(was_training)
