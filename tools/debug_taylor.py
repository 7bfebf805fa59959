"""debug: TaylorDecompositionRule on an nn.Linear whose class forward was replaced by monkey_patch(vision_transformer)"""
import sys, warnings
sys.path.insert(0, "lrp-explains-transformers_b200"); sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import torch
from torchvision.models import vision_transformer
from lxt_b200.efficient import monkey_patch
import lxt_b200.explicit.rules as rules
from oracle import attnlrp_oracle as O
from helpers import rel_l2
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    monkey_patch(vision_transformer)
g = torch.Generator().manual_seed(4)
x, W, R = torch.rand(6, 16, generator=g) + 0.5, torch.rand(8, 16, generator=g) + 0.5, torch.randn(6, 8, generator=g)
lin = torch.nn.Linear(16, 8, bias=False).cuda()
lin.weight.data.copy_(W)
lin.weight.requires_grad_(False)
xs = x.cuda().requires_grad_()
y = lin(xs)
print("fwd err", rel_l2(y.detach().cpu(), x @ W.t()))
y.backward(R.cuda())
print("bwd err", rel_l2(xs.grad.cpu(), R @ W))
rule = rules.TaylorDecompositionRule(lin, ref=(torch.zeros(6, 16, device="cuda"),), bias=False)
xs = x.cuda().requires_grad_()
y = rule(xs)
y.backward(R.cuda())
exp = O.linear_epsilon_relevance(x, W, None, R, 1e-6)
print("taylor err", rel_l2(xs.grad.cpu(), exp))
