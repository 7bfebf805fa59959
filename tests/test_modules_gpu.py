"""GPU: the module layer of the explicit API (`lxt_b200.explicit.modules`, `.special`) against golden vectors produced by the REAL
reference's `lxt.explicit.modules` (tests/golden/make_golden_modules.py -> explicit_modules.npz; fp32, CPU).
Tolerances: element-wise fp32 kernels 1e-5; rules whose normalised relevance is a bf16 GEMM operand 2.5e-3; fp32 attention path 1e-4."""
import pytest
import torch
import torch.nn as nn

from helpers import load_npz, rel_l2

pytestmark = pytest.mark.gpu
T = torch.from_numpy


@pytest.fixture(scope="module")
def Z():
    return {k: T(v) for k, v in load_npz("explicit_modules.npz").items()}


def _run(mod, x, seed, **kw):
    xi = x.cuda().requires_grad_()
    y = mod(xi, **kw) if not isinstance(mod, tuple) else None
    y0 = y[0] if isinstance(y, tuple) else y
    y0.backward(seed.cuda().to(y0.dtype))
    return y, xi.grad.float().cpu()


def test_rule_modules_match_the_reference(Z):
    import lxt_b200.explicit.modules as lm
    y, r = _run(lm.SoftmaxDT(dim=-1, temperature=2.0), Z["x"], Z["R"])
    assert rel_l2(y.detach().cpu(), Z["softmax_y"]) < 1e-5 and rel_l2(r, Z["softmax_R"]) < 1e-5

    lin = nn.Linear(128, 64)
    with torch.no_grad():
        lin.weight.copy_(Z["lin_w"]); lin.bias.copy_(Z["lin_b"])
    le = lm.initialize_bias(lin.cuda(), lm.LinearEpsilon)
    assert isinstance(le, lm.LinearEpsilon) and le.weight is lin.weight and le.epsilon == 1e-6
    y, r = _run(le, Z["lin_x"], Z["lin_Rout"])
    assert rel_l2(y.detach().cpu(), Z["lin_y"]) < 2.5e-3 and rel_l2(r, Z["lin_R"]) < 2.5e-3

    rms = lm.RMSNormIdentity(128, eps=1e-5).cuda()
    with torch.no_grad():
        rms.weight.copy_(Z["rms_w"])
    y, r = _run(rms, Z["x"], Z["R"])
    assert rel_l2(y.detach().cpu(), Z["rms_y"]) < 1e-5 and rel_l2(r, Z["rms_R"]) < 1e-5

    ln = nn.LayerNorm(128)
    with torch.no_grad():
        ln.weight.copy_(Z["ln_w"]); ln.bias.copy_(Z["ln_b"])
    lne = lm.initialize_bias(ln.cuda(), lm.LayerNormEpsilon)
    y, r = _run(lne, Z["x"], Z["R"])
    assert rel_l2(y.detach().cpu(), Z["ln_y"]) < 1e-5 and rel_l2(r, Z["ln_R"]) < 1e-3


def _cp_module(Z, dtype=torch.float32):
    import lxt_b200.explicit.modules as lm
    mha = nn.MultiheadAttention(128, 2, batch_first=True)
    with torch.no_grad():
        mha.in_proj_weight.copy_(Z["mha_in_w"]); mha.in_proj_bias.copy_(Z["mha_in_b"])
        mha.out_proj.weight.copy_(Z["mha_out_w"]); mha.out_proj.bias.copy_(Z["mha_out_b"])
    for p in mha.parameters():
        p.requires_grad_(False)
    return lm.initialize_MHA(mha.cuda().to(dtype), lm.MultiheadAttention_CP)


def test_multihead_attention_cp_matches_the_reference(Z):
    from lxt_b200 import ops
    cp = _cp_module(Z)
    assert cp.head_dim == 64 and cp.num_heads == 2 and cp.batch_first is True and cp.q_proj_weight.shape == (128, 128)

    def run(**kw):
        xi = Z["x"].cuda().requires_grad_()
        n0 = ops.launch_count()
        out, w = cp(xi, xi, xi, **kw)
        out.backward(Z["R"].cuda())
        return out.detach().cpu(), w, xi.grad.cpu(), ops.launch_count() - n0

    y, w, r, n = run(need_weights=False)
    assert w is None and n >= 8                      # projections, flash forward / backward, epsilon division, products: all kernels
    assert rel_l2(y, Z["mha_y_nw"]) < 1e-4 and rel_l2(r, Z["mha_R_nw"]) < 1e-4
    y, w, r, _ = run(need_weights=True)
    assert rel_l2(y, Z["mha_y"]) < 1e-4 and rel_l2(r, Z["mha_R"]) < 1e-4 and rel_l2(w.cpu(), Z["mha_w"]) < 1e-4
    y, w, r, _ = run(need_weights=False, key_padding_mask=Z["mha_kpm"].cuda())
    assert rel_l2(y, Z["mha_y_kpm"]) < 1e-4 and rel_l2(r, Z["mha_R_kpm"]) < 1e-4
    with pytest.raises(NotImplementedError):
        run(need_weights=False, attn_mask=torch.zeros(17, 17, device="cuda"))
    holes = torch.zeros(2, 17, dtype=torch.bool, device="cuda")
    holes[:, 3] = True
    with pytest.raises(NotImplementedError):
        run(need_weights=False, key_padding_mask=holes)


def test_multihead_attention_cp_bf16(Z):
    """bf16 operands: the forward on the tcgen05 flash kernel; the relevance is only checked to be finite and of the reference's
    magnitude — R / (Y + 1e-6) on a bf16 Y is as ill-conditioned as the rule itself"""
    cp = _cp_module(Z, torch.bfloat16)
    xi = Z["x"].cuda().to(torch.bfloat16).requires_grad_()
    out, _ = cp(xi, xi, xi, need_weights=False)
    out.backward(Z["R"].cuda().to(torch.bfloat16))
    assert rel_l2(out.detach().float().cpu(), Z["mha_y_nw"]) < 3e-2
    assert torch.isfinite(xi.grad).all() and rel_l2(xi.grad.float().cpu(), Z["mha_R_nw"]) < 0.2
