"""CPU: host-side logic of the drop-in API (patch plumbing, registries, sharding) and the loud-failure contract
(no CPU fallback: rules / ops raise on CPU tensors instead of silently computing elsewhere)."""
import os
import types
import warnings

import pytest
import torch

from lxt_b200 import _capi, ops
from lxt_b200.efficient import monkey_patch
from lxt_b200.efficient import patches as P
from lxt_b200.efficient import rules as R
from lxt_b200.efficient.models import DEFAULT_MAP, get_default_map


def test_patch_method_and_keep_original():
    class A(torch.nn.Identity):  # forward lives in torch.nn.modules.linear, i.e. "not yet patched"
        pass

    def fwd(self, x):
        return x * 2

    orig = A.forward
    assert P.patch_method(fwd, A, keep_original=True) is True
    assert A.forward is fwd and A.original_forward is orig
    assert A().forward(3) == 6


def test_already_patched_guard_is_module_string_equality():
    # reference lxt/efficient/patches.py:40: functions from the same module count as "already patched"
    class B:
        forward = P.dropout_forward

    with pytest.warns(UserWarning):
        assert P.patch_method(P.rms_norm_forward, B) is False


def test_replace_module_copies_public_attributes():
    src, dst = types.ModuleType("src"), types.ModuleType("dst")
    src.a, src._b = 1, 2
    assert P.replace_module(src, dst) is True and dst.a == 1 and dst._b == 2
    assert P.replace_module(src, src) is False


def test_default_map_and_unknown_module():
    from transformers.models.llama import modeling_llama
    assert modeling_llama in DEFAULT_MAP and "attnLRP" in dir(__import__("lxt_b200.efficient.models.llama", fromlist=["x"]))
    with pytest.raises(ValueError, match="not yet supported"):
        get_default_map(types)


def test_monkey_patch_reports_failed_patchers():
    calls = []

    class T1:
        pass

    class T2:
        pass

    pm = {T1: lambda t: calls.append(t) or True, T2: lambda t: calls.append(t) or False}
    with pytest.warns(UserWarning, match="Failed to patch T2"):
        monkey_patch(types, patch_map=pm)
    assert calls == [T1, T2]


def test_attention_wrapper_shape_of_api():
    w = P.wrap_attention_forward(lambda *a, **k: None)
    assert w.__name__ == "attention_forward" and w.__module__ == P.__name__
    cp = P.cp_wrap_attention_forward(lambda *a, **k: None)
    assert cp.__name__ == "cp_attention_forward"


def test_rules_fail_loudly_on_cpu_tensors():
    x = torch.randn(4, 8, requires_grad=True)
    assert R.stop_gradient(x).requires_grad is False
    y = R.divide_gradient(x, 4)  # forward is the identity and launches nothing
    with pytest.raises(_capi.LrpError):
        y.sum().backward()
    with pytest.raises(_capi.LrpError):
        R.identity_rule_implicit(torch.nn.functional.silu, x).sum().backward()
    with pytest.raises(_capi.LrpError):
        ops.eps_div(x.detach(), x.detach(), 1e-6)
    with pytest.raises(_capi.LrpError):
        ops.gemm(torch.zeros(8, 8, dtype=torch.bfloat16), torch.zeros(8, 8, dtype=torch.bfloat16), torch.zeros(8, 8), b_layout=0)


def test_engine_refuses_cpu():
    from lxt_b200.engine import LlamaAttnLRPEngine, LlamaDims
    with pytest.raises(RuntimeError):
        LlamaAttnLRPEngine(LlamaDims(d=64, I=128, H=1, Hkv=1, D=64, L=1, V=16), torch.device("cpu"), {})


def test_shard_range_is_balanced_and_contiguous():
    from lxt_b200.dist import shard_range
    for n in (1, 7, 32, 33):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_conservation_check_flag_and_wrapper():
    """lxt/explicit/check.py:6-15 + functional.py:12-37: uniform redistribution that conserves the relevance sum"""
    from lxt_b200.explicit import functional as lf
    from lxt_b200.explicit.check import conservation_check

    class Ctx:
        pass

    @lf.conservation_check_wrap
    def backward(ctx, *r):
        return (r[0] * 2.0, None, r[0][:, :2] * 3.0)

    R = torch.arange(12.0).reshape(3, 4)
    out = backward(Ctx(), R)
    assert torch.equal(out[0], R * 2) and out[1] is None
    with conservation_check():
        assert lf.CONSERVATION_CHECK_FLAG[0] is True
        u = backward(Ctx(), R)
        assert u[1] is None and u[0].shape == (3, 4) and u[2].shape == (3, 2)
        assert abs(float(u[0].sum() + u[2].sum()) - float(R.sum())) < 1e-4     # conserved
        assert float(u[0].max() - u[0].min()) == 0.0                               # uniform
    assert lf.CONSERVATION_CHECK_FLAG[0] is False


def test_install_as_lxt_registers_the_reference_module_names():
    """user scripts written against the reference import `lxt.efficient`, `lxt.explicit.functional`, ... and, for the ViT recipe,
    `zennit.rules.Gamma` / `zennit.composites.LayerMapComposite` (examples/vit_torch.py:6-10)"""
    import subprocess
    import sys
    code = (
        "import sys; sys.path.insert(0, 'lrp-explains-transformers_b200')\n"
        "import lxt_b200; lxt_b200.install_as_lxt()\n"
        "import lxt.explicit.functional as lf, lxt.explicit.modules as lm, lxt.explicit.special as ls, lxt.explicit.rules as lr\n"
        "from lxt.efficient import monkey_patch, monkey_patch_zennit\n"
        "import lxt_b200.explicit.functional as own\n"
        "assert lf is own and lm.LinearEpsilon.__module__ == 'lxt_b200.explicit.modules'\n"
        "import zennit.rules as zr; from zennit.composites import LayerMapComposite\n"
        "assert zr.Gamma(0.25).gamma == 0.25 and callable(LayerMapComposite([]).register)\n"
        "print('ok')\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], cwd=root, capture_output=True, text=True)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]


def test_explicit_module_initialisers_share_tensors_and_geometry():
    """lxt.explicit.modules initialisers (reference modules.py:127-214): replacements are built from the original's constructor
    arguments, share its parameters, and the packed in-projection of nn.MultiheadAttention becomes three views"""
    import torch.nn as nn
    import lxt_b200.explicit.modules as lm
    E = 128
    mha = nn.MultiheadAttention(E, 2, batch_first=True)
    cp = lm.initialize_MHA(mha, lm.MultiheadAttention_CP)
    assert (cp.embed_dim, cp.num_heads, cp.head_dim, cp.batch_first) == (E, 2, 64, True)
    assert torch.equal(cp.q_proj_weight, mha.in_proj_weight[:E]) and torch.equal(cp.k_proj_weight, mha.in_proj_weight[E:2 * E])
    assert torch.equal(cp.v_proj.weight, mha.in_proj_weight[2 * E:]) and torch.equal(cp.bias_k, mha.in_proj_bias[E:2 * E])
    assert cp.q_proj_weight.data_ptr() == mha.in_proj_weight.data_ptr() and cp.out_proj.weight is mha.out_proj.weight
    cp2 = lm.initialize_MHA(nn.MultiheadAttention(E, 2, kdim=64, vdim=32), lm.MultiheadAttention_CP)
    assert cp2.k_proj_weight.shape == (E, 64) and cp2.v_proj.weight.shape == (E, 32) and cp2.batch_first is False
    with pytest.raises(NotImplementedError):
        lm.initialize_MHA(nn.MultiheadAttention(E, 2, add_bias_kv=True), lm.MultiheadAttention_CP)
    lin = nn.Linear(8, 4, bias=False)
    le = lm.initialize_bias(lin, lm.LinearEpsilon)
    assert isinstance(le, lm.LinearEpsilon) and le.bias is None and le.weight is lin.weight and le.epsilon == 1e-6
    ln = nn.LayerNorm(16)
    l2 = lm.initialize_bias(ln, lm.LayerNormEpsilon)
    assert l2.weight is ln.weight and l2.bias is ln.bias and l2.eps == ln.eps
    sm = lm.initialize_generic(nn.Softmax(dim=-1), lm.SoftmaxDT)
    assert sm.dim == -1 and sm.temperature == 1.0 and sm.dtype is None
    assert set(lm.INIT_MODULE_MAPPING) == {lm.SoftmaxDT, lm.LinearEpsilon, lm.RMSNormIdentity, lm.LayerNormEpsilon, lm.MultiheadAttention_CP}
