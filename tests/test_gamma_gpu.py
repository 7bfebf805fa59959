"""GPU: the Gamma rule of the reference's ViT recipe (zennit Gamma in Gradient x Input space, lxt/efficient/zennit_patches.py:26-62,
examples/vit_torch.py:59-65) on the B200 kernels against oracle/gamma_oracle.py (PARITY UNPINNED: zennit absent, see its header).
Tolerances: fp32 tensors (split-bf16 GEMMs) rel-L2 <= 1e-3 against float64; bf16 tensors <= 3e-2 (s = R / z is a bf16 GEMM operand)."""
import copy
import warnings

import pytest
import torch
import torch.nn as nn

from helpers import load_npz, rel_l2
from oracle import gamma_oracle as GO

pytestmark = pytest.mark.gpu


def _run_linear(x, W, b, gy, gamma, dtype):
    from lxt_b200.efficient.zennit_rules import Gamma, LayerMapComposite
    lin = nn.Linear(W.shape[1], W.shape[0], bias=b is not None).cuda().to(dtype)
    with torch.no_grad():
        lin.weight.copy_(W)
        if b is not None:
            lin.bias.copy_(b)
    comp = LayerMapComposite([(nn.Linear, Gamma(gamma))])
    comp.register(lin)
    xg = x.cuda().to(dtype).requires_grad_()
    y = lin(xg)
    y.backward(gy.cuda().to(dtype))
    comp.remove()
    return y.detach().float().cpu(), xg.grad.float().cpu()


@pytest.mark.parametrize("gamma", [0.25, 100.0])
def test_gamma_linear_vit_mlp_shape(gamma):
    g = torch.Generator().manual_seed(4)
    T, K, N = 2 * 197, 1024, 4096
    x = torch.randn(T, K, generator=g)
    W = torch.randn(N, K, generator=g) * 0.03
    b = torch.randn(N, generator=g) * 0.1
    gy = torch.randn(T, N, generator=g)
    want = GO.gamma_linear_gxi(x.double(), W.double(), b.double(), gy.double(), gamma).float()
    y, got = _run_linear(x, W, b, gy, gamma, torch.float32)
    assert rel_l2(y, (x.double() @ W.double().t() + b.double()).float()) < 1e-5
    assert rel_l2(got, want) < 1e-3
    xb, Wb, bb, gb = (t.bfloat16().float() for t in (x, W, b, gy))
    want16 = GO.gamma_linear_gxi(xb.double(), Wb.double(), bb.double(), gb.double(), gamma).float()
    _, got16 = _run_linear(xb, Wb, bb, gb, gamma, torch.bfloat16)
    assert rel_l2(got16, want16) < 3e-2


def test_gamma_linear_without_bias_conserves_relevance():
    g = torch.Generator().manual_seed(5)
    x = torch.randn(64, 256, generator=g)
    W = torch.randn(128, 256, generator=g) * 0.1
    gy = torch.randn(64, 128, generator=g)
    y, gx = _run_linear(x, W, None, gy, 0.25, torch.float32)
    assert abs(float((x * gx).sum()) - float((gy * y).sum())) < 1e-3 * float((gy * y).abs().sum())


def test_gamma_patch_embedding_conv():
    from lxt_b200.efficient.zennit_rules import Gamma, LayerMapComposite
    g = torch.Generator().manual_seed(6)
    x = torch.randn(2, 3, 64, 64, generator=g)
    conv = nn.Conv2d(3, 128, 16, 16)
    gy = torch.randn(2, 128, 4, 4, generator=g)
    want = GO.gamma_conv2d_gxi(x.double(), conv.weight.detach().double(), conv.bias.detach().double(), gy.double(), 16, 100.0).float()
    cc = copy.deepcopy(conv).cuda()
    comp = LayerMapComposite([(nn.Conv2d, Gamma(100.0))])
    comp.register(cc)
    xg = x.cuda().requires_grad_()
    cc(xg).backward(gy.cuda())
    assert rel_l2(xg.grad.cpu(), want) < 1e-3
    with pytest.raises(NotImplementedError):
        bad = nn.Conv2d(3, 8, 3, 1, padding=1).cuda()
        LayerMapComposite([(nn.Conv2d, Gamma(1.0))]).register(bad)
        bad(torch.randn(1, 3, 8, 8, device="cuda"))


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-3), (torch.bfloat16, 6e-2)])
def test_vit_recipe_with_gamma_matches_the_cpu_restatement(dtype, tol):
    """examples/vit_torch.py:15-16,59-65,84-91 end to end on a small torchvision ViT: monkey_patch(vision_transformer) + Gamma
    composite, heat-map (x * x.grad).sum(1)."""
    from torchvision.models import vision_transformer
    from lxt_b200.efficient import monkey_patch, monkey_patch_zennit
    from lxt_b200.efficient.zennit_rules import Gamma, LayerMapComposite
    z = load_npz("vit_tiny.npz")
    base = vision_transformer.VisionTransformer(image_size=64, patch_size=16, num_layers=2, num_heads=2, hidden_dim=128, mlp_dim=256,
                                                num_classes=16)
    base.load_state_dict({k[3:]: torch.from_numpy(v) for k, v in z.items() if k.startswith("sd_")})
    for p in base.parameters():
        p.requires_grad_(False)
    base.eval()
    x = torch.from_numpy(z["x"])
    cls = torch.from_numpy(z["cls"])
    conv_gamma, lin_gamma = 100.0, 0.25
    # CPU restatement, float64 (for the bf16 run: on the bf16-rounded weights and pixels)
    ref_model = copy.deepcopy(base)
    xin = x
    if dtype == torch.bfloat16:
        ref_model = ref_model.to(torch.bfloat16).float()
        xin = x.bfloat16().float()
    ref_model = GO.patch_vit_cpu(ref_model.double(), conv_gamma, lin_gamma)
    xr = xin.double().requires_grad_()
    ref_model(xr)[torch.arange(2), cls].sum().backward()
    want = (xr * xr.grad).sum(1).float()
    # B200 path: same user code as the reference example.  monkey_patch replaces CLASS methods process-wide (as in the reference):
    # they are put back afterwards so that later test modules see stock torch.nn again.
    patched = (nn.GELU, nn.LayerNorm, nn.MultiheadAttention, nn.Linear)
    saved = [(c, c.forward, c.__dict__.get("original_forward")) for c in patched]
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            monkey_patch(vision_transformer)
        monkey_patch_zennit()
        model = copy.deepcopy(base).cuda().to(dtype)
        comp = LayerMapComposite([(nn.Conv2d, Gamma(conv_gamma)), (nn.Linear, Gamma(lin_gamma))])
        comp.register(model)
        xg = xin.cuda().to(dtype).requires_grad_()
        model(xg)[torch.arange(2), cls.cuda()].sum().backward()
        comp.remove()
    finally:
        for c, fwd, orig in saved:
            c.forward = fwd
            if orig is None and "original_forward" in c.__dict__:
                delattr(c, "original_forward")
    heat = (xg * xg.grad).float().sum(1).cpu()
    err = rel_l2(heat, want)
    print(f"ViT + Gamma heat-map ({dtype}) rel-L2 vs CPU restatement: {err:.3e}")
    assert err < tol
    # the rule changes the explanation: it must differ from the plain-gradient heat-map of the golden
    assert rel_l2(want, torch.from_numpy(z["heat"])) > 1e-2
