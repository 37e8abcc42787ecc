"""TransMIL aggregator (stamp_b200/transmil.py) against the reference's own module: tests/golden/transmil.npz holds the
logits of the reference TransMIL (imported by path in oracle/make_golden_transmil.py) for seeded weights and bags of 300 and
1100 tiles; weights and inputs are regenerated here from the same seeds (inputs verified by checksum)."""

from pathlib import Path

import numpy as np
import pytest
import torch

GOLD = Path(__file__).parent / "golden" / "transmil.npz"


def _bags(n: int, g: torch.Generator) -> torch.Tensor:
    return torch.randn(2, n, 64, generator=g).half().float()


def test_transmil_state_dict_is_the_reference_one():
    from oracle.transmil_weights import transmil_state_dict
    from stamp_b200.transmil import TransMIL

    model = TransMIL(dim_output=3, dim_input=64, dim_hidden=512)
    sd = transmil_state_dict(3, 64, 512)
    assert set(model.state_dict()) == set(sd)
    model.load_state_dict(sd, strict=True)
    with pytest.raises((RuntimeError, NotImplementedError)):
        model.eval()(torch.zeros(1, 4, 64))


@pytest.mark.gpu
def test_transmil_matches_reference_golden(cuda_device):
    from oracle.transmil_weights import transmil_state_dict
    from stamp_b200.transmil import TransMIL

    z = np.load(GOLD)
    model = TransMIL(dim_output=3, dim_input=64, dim_hidden=512)
    model.load_state_dict(transmil_state_dict(3, 64, 512), strict=True)
    model = model.to(cuda_device).eval()
    g = torch.Generator().manual_seed(9)
    for n in (300, 1100):
        bags = _bags(n, g)
        assert abs(bags.double().sum().item() - float(z[f"bags_checksum_{n}"])) < 1e-9      # the golden's inputs
        with torch.inference_mode():
            got = model(bags.to(cuda_device)).float().cpu()
        want = torch.from_numpy(z[f"logits_{n}"])
        err = ((got - want).norm(dim=1) / want.norm(dim=1)).max().item()
        print(f"TransMIL, {n} tiles (_fc1 on the tcgen05 GEMM, fp16 operands): max per-bag relative error {err:.2e}")
        assert got.shape == (2, 3) and err < 1e-3, (n, err, got, want)
        model.fc1_fp32 = True                        # everything in fp32: the restatement itself is exact
        with torch.inference_mode():
            got32 = model(bags.to(cuda_device)).float().cpu()
        model.fc1_fp32 = False
        err32 = ((got32 - want).norm(dim=1) / want.norm(dim=1)).max().item()
        print(f"TransMIL, {n} tiles, all fp32: {err32:.2e}")
        assert err32 < 1e-4, (n, err32)
        # the reference's batch-wide scaling of the pseudo-inverse start: alone, bag 1 comes out differently
        with torch.inference_mode():
            alone = model(bags[1:].to(cuda_device)).float().cpu()
        assert (alone - got[1:]).abs().max() > 1e-5 or n == 300


@pytest.mark.gpu
@pytest.mark.parametrize("M,N,K,batch", [(70, 33, 19, 3), (256, 256, 256, 8), (5, 130, 64, 1), (129, 64, 300, 2)])
def test_sgemm_batched_f32_all_modes(cuda_device, M, N, K, batch):
    """stamp_sgemm_batched_f32 at ragged shapes: A B^T and A B, the (eye * I - B) operand of the Newton-Schulz steps, bias,
    ReLU and accumulation, against fp64."""
    from stamp_b200 import _lib
    from stamp_b200.transmil import _bind

    lib = _bind()
    g = torch.Generator().manual_seed(M * 1000 + N)
    st = torch.cuda.current_stream().cuda_stream
    A = torch.randn(batch, M, K, generator=g).to(cuda_device)
    bias = torch.randn(N, generator=g).to(cuda_device)
    for trans_b, eye, mode, use_bias in ((1, 0.0, 0, True), (0, 0.0, 0, False), (1, 0.0, 3, True), (0, 0.0, 1, True)) + \
            (((0, 7.0, 0, False),) if N == K else ()):
        B = torch.randn(batch, *((N, K) if trans_b else (K, N)), generator=g).to(cuda_device)
        C0 = torch.randn(batch, M, N, generator=g).to(cuda_device)
        C = C0.clone()
        _lib.check(lib.stamp_sgemm_batched_f32(A.data_ptr(), K, M * K, B.data_ptr(), B.shape[2], B.shape[1] * B.shape[2], C.data_ptr(), N,
                                               M * N, M, N, K, batch, trans_b, 0.5, eye, bias.data_ptr() if use_bias else None, mode, st),
                   "sgemm")
        Bd = B.double().transpose(1, 2) if trans_b else B.double()
        if eye:
            Bd = eye * torch.eye(N, device=cuda_device, dtype=torch.float64) - Bd
        want = 0.5 * A.double() @ Bd + (bias.double() if use_bias else 0.0)
        if mode & 2:
            want = want.clamp_min(0.0)
        if mode & 1:
            want = want + C0.double()
        assert ((C.double() - want).norm() / want.norm()).item() < 1e-6, (trans_b, eye, mode)


@pytest.mark.gpu
@pytest.mark.parametrize("nq,nk,splits", [(256, 4352, 17), (100, 77, 1), (64, 1000, 5), (300, 256, 1), (7, 33, 9)])
def test_attention_f32_key_splits(cuda_device, nq, nk, splits):
    """stamp_attention_f32 with the keys shared among several CTAs equals softmax(scale Q K^T) V in fp64."""
    from stamp_b200 import _lib
    from stamp_b200.transmil import _bind

    lib = _bind()
    H = 2
    g = torch.Generator().manual_seed(nq + nk)
    q, k, v = (torch.randn(n, H * 64, generator=g).to(cuda_device) for n in (nq, nk, nk))
    out = torch.empty(nq, H * 64, device=cuda_device)
    part = torch.empty(splits * H * nq * 66, device=cuda_device)
    _lib.check(lib.stamp_attention_f32(q.data_ptr(), H * 64, k.data_ptr(), H * 64, v.data_ptr(), H * 64, out.data_ptr(), H * 64, nq, nk,
                                       H, 0.125, part.data_ptr(), splits, torch.cuda.current_stream().cuda_stream), "attention_f32")
    sp = lambda t: t.double().reshape(-1, H, 64).transpose(0, 1)      # noqa: E731
    want = ((sp(q) @ sp(k).transpose(1, 2) * 0.125).softmax(-1) @ sp(v)).transpose(0, 1).reshape(nq, H * 64)
    assert ((out.double() - want).norm() / want.norm()).item() < 1e-6


@pytest.mark.gpu
def test_reference_unit_test_shapes(cuda_device):
    """The reference's own model tests (tests/test_model.py:77-166): MLP on [7, 33] vectors, TransMIL on 7 bags of 76 tiles
    with 457 input features (not a multiple of the GEMM's K granularity: runs zero-padded), coords / mask keywords accepted
    and ignored, two forwards identical; the TransMIL logits against the reference module's (golden `logits_odd`)."""
    from oracle.transmil_weights import transmil_state_dict
    from stamp_b200.mlp import MLP
    from stamp_b200.transmil import TransMIL

    z = np.load(GOLD)
    g = torch.Generator().manual_seed(9)
    for n in (300, 1100):
        _bags(n, g)                                                   # advance the generator as the golden's script did
    bags = torch.rand(7, 76, 457, generator=g).half().float()
    assert abs(bags.double().sum().item() - float(z["bags_checksum_odd"])) < 1e-9
    model = TransMIL(dim_output=4, dim_input=457, dim_hidden=512)
    model.load_state_dict(transmil_state_dict(4, 457, 512), strict=True)
    model = model.to(cuda_device).eval()
    mask = torch.arange(76)[None, :].repeat(7, 1) >= torch.randint(1, 76, (7, 1))
    with torch.inference_mode():
        a = model.forward(bags.to(cuda_device), coords=torch.rand(7, 76, 2, device=cuda_device), mask=mask.to(cuda_device))
        b = model.forward(bags.to(cuda_device), coords=torch.rand(7, 76, 2, device=cuda_device), mask=mask.to(cuda_device))
    assert a.shape == (7, 4) and torch.equal(a, b)
    want = torch.from_numpy(z["logits_odd"])
    err = ((a.float().cpu() - want).norm(dim=1) / want.norm(dim=1)).max().item()
    model.fc1_fp32 = True
    with torch.inference_mode():
        a32 = model.forward(bags.to(cuda_device))
    model.fc1_fp32 = False
    err32 = ((a32.float().cpu() - want).norm(dim=1) / want.norm(dim=1)).max().item()
    print(f"TransMIL, reference test shape (76 tiles = one token per landmark, 457 features): max per-bag relative error "
          f"{err:.2e} (_fc1 fp16 operands), {err32:.2e} (all fp32)")
    # one token per landmark: the pseudo-inverse is of the full 256 x 256 attention matrix, the worst-conditioned case
    assert err32 < 1e-4 and err < 2e-3, (err, err32)              # measured 2.9e-6 / 5.2e-4

    mlp = MLP(dim_output=4, dim_input=33, dim_hidden=64, num_layers=3, dropout=0.1).to(cuda_device).eval()
    feats = torch.rand(7, 33, device=cuda_device)
    with torch.inference_mode():
        l1, l2 = mlp.forward(feats), mlp.forward(feats)
        want = mlp.mlp(feats)                                          # the same nn.Sequential in eager torch on the GPU
    assert l1.shape == (7, 4) and torch.equal(l1, l2)
    assert ((l1 - want).norm() / want.norm()).item() < 1e-5            # eager torch may use TF32-free fp32: same arithmetic
