"""CPU: the closed-form backward restatements (oracle/backward_ref.py, the checkers of next round's backward kernels) against
torch autograd through the forward formulas of oracle/model_ref.py."""
import torch

from oracle import backward_ref as BR
from oracle import model_ref as R


def _g(seed):
    return torch.Generator().manual_seed(seed)


def test_rmsnorm_and_sandwich_bwd():
    g = _g(0)
    x = torch.randn(5, 7, 64, generator=g).requires_grad_(True)
    w = (torch.randn(64, generator=g) * 0.1).requires_grad_(True)
    dy = torch.randn(5, 7, 64, generator=g)
    R._rms(x, w, 1e-6).backward(dy)
    dx, dw = BR.rmsnorm_bwd(x.detach(), w.detach(), dy, 1e-6)
    assert (dx - x.grad).abs().max() < 1e-5 and (dw - w.grad).abs().max() < 1e-4
    xr = torch.randn(3, 64, generator=g).requires_grad_(True)
    br = torch.randn(3, 64, generator=g).requires_grad_(True)
    w2 = (torch.randn(64, generator=g) * 0.1).requires_grad_(True)
    d_out = torch.randn(3, 64, generator=g)
    (xr + R._rms(br, w2, 1e-6)).backward(d_out)
    dxr, dbr, dw2 = BR.sandwich_bwd(xr.detach(), br.detach(), w2.detach(), d_out, 1e-6)
    assert torch.equal(dxr, xr.grad) and (dbr - br.grad).abs().max() < 1e-5 and (dw2 - w2.grad).abs().max() < 1e-5


def test_geglu_bwd():
    g = _g(1)
    gate = (torch.randn(4, 96, generator=g) * 2).requires_grad_(True)
    up = torch.randn(4, 96, generator=g).requires_grad_(True)
    da = torch.randn(4, 96, generator=g)
    (R.gelu_tanh(gate) * up).backward(da)
    dg, du = BR.geglu_bwd(gate.detach(), up.detach(), da)
    assert (dg - gate.grad).abs().max() < 1e-5 and (du - up.grad).abs().max() < 1e-6


def test_rope_bwd_is_the_inverse_rotation():
    g = _g(2)
    for pos in (torch.arange(1, 12), torch.stack([torch.arange(1, 12), torch.arange(3, 14)])):
        B = 1 if pos.dim() == 1 else 2
        x = torch.randn(B, 3, 11, 32, generator=g).requires_grad_(True)
        dy = torch.randn(B, 3, 11, 32, generator=g)
        R._rope(x, pos, 10000.0).backward(dy)
        dx = BR.rope_bwd(dy, pos, 10000.0)
        assert (dx - x.grad).abs().max() < 1e-5
        assert (R._rope(dx, pos, 10000.0) - dy).abs().max() < 1e-5        # orthogonal: forward(backward(dy)) = dy


def test_softcap_attention_bwd_all_training_masks():
    g = _g(3)
    B, H, S, D, P = 2, 3, 19, 16, 11
    for name, mask in (("bidirectional", None), ("causal", torch.arange(S)[None, :] > torch.arange(S)[:, None]),
                       ("prefix_lm", torch.arange(S)[None, :] > torch.clamp(torch.arange(S)[:, None], min=P - 1))):
        for cap in (50.0, 0.0):
            q, k, v = (torch.randn(B, H, S, D, generator=g).requires_grad_(True) for _ in range(3))
            do = torch.randn(B, H, S, D, generator=g)
            s = (q @ k.transpose(2, 3)) * 0.25
            if cap:
                s = torch.tanh(s / cap) * cap
            if mask is not None:
                s = s.masked_fill(mask, float("-inf"))
            (torch.softmax(s, -1) @ v).backward(do)
            dq, dk, dv = BR.softcap_attention_bwd(q.detach(), k.detach(), v.detach(), do, 0.25, cap, mask)
            for a, b, nm in ((dq, q.grad, "dq"), (dk, k.grad, "dk"), (dv, v.grad, "dv")):
                assert (a - b).abs().max() < 2e-5, (name, cap, nm)


def test_lora_linear_bwd_never_forms_the_base_weight_gradient():
    g = _g(4)
    x = torch.randn(6, 5, 48, generator=g).requires_grad_(True)
    w = torch.randn(40, 48, generator=g) * 0.1
    A = (torch.randn(4, 48, generator=g) * 0.2).requires_grad_(True)
    Bm = (torch.randn(40, 4, generator=g) * 0.2).requires_grad_(True)
    dy = torch.randn(6, 5, 40, generator=g)
    s = 2.0
    (torch.nn.functional.linear(x, w) + s * torch.nn.functional.linear(torch.nn.functional.linear(x, A), Bm)).backward(dy)
    dx, gA, gB = BR.lora_linear_bwd(x.detach(), dy, w, A.detach(), Bm.detach(), s)
    assert (dx - x.grad).abs().max() < 1e-5 and (gA - A.grad).abs().max() < 1e-4 and (gB - Bm.grad).abs().max() < 1e-4
