"""The reference's ViT recipe (examples/vit_torch.py) with only its imports redirected: `install_as_lxt()` registers this package
under the name `lxt` and, since `zennit` is not installed, stand-ins for `zennit.rules.Gamma` / `zennit.composites.LayerMapComposite`
backed by the B200 Gamma-rule kernels.  Everything below the imports is the reference's user code (random weights and a synthetic
image instead of the downloaded checkpoint and the picture of a dog).

    python examples/vit_gamma_recipe.py        # needs a B200 and the built liblrp_b200.so
"""
import itertools
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lrp-explains-transformers_b200"))
import lxt_b200  # noqa: E402

lxt_b200.install_as_lxt()

import torch  # noqa: E402
from torchvision.models import vision_transformer  # noqa: E402
from zennit.composites import LayerMapComposite  # noqa: E402   (the stand-in when zennit is absent)
import zennit.rules as z_rules  # noqa: E402
from lxt.efficient import monkey_patch, monkey_patch_zennit  # noqa: E402

monkey_patch(vision_transformer, verbose=True)
monkey_patch_zennit(verbose=True)


def main():
    torch.manual_seed(0)
    model = vision_transformer.vit_b_16(weights=None)
    torch.nn.init.normal_(model.heads.head.weight, std=0.02)          # torchvision zero-fills the head
    model = model.eval().cuda().to(torch.bfloat16)
    for p in model.parameters():
        p.requires_grad_(False)
    image = torch.randn(1, 3, 224, 224, generator=torch.Generator().manual_seed(1)).cuda().to(torch.bfloat16)

    for conv_gamma, lin_gamma in itertools.product([0.1, 0.25, 100], [0, 0.01, 0.05, 0.1, 1]):
        comp = LayerMapComposite([(torch.nn.Conv2d, z_rules.Gamma(conv_gamma)), (torch.nn.Linear, z_rules.Gamma(lin_gamma))])
        comp.register(model)
        x = image.clone().requires_grad_()
        y = model(x)
        y[0, y[0].argmax()].backward()
        heatmap = (x * x.grad).float().sum(1)
        comp.remove()
        print(f"Gamma conv {conv_gamma:>6} linear {lin_gamma:>5}: heat-map sum {float(heatmap.sum()):+.4e}, "
              f"max |R| {float(heatmap.abs().max()):.3e}, finite {bool(torch.isfinite(heatmap).all())}")


if __name__ == "__main__":
    main()
