"""BASELINE configs[3]: torchvision ViT-L/16, random init, 224x224 synthetic images, pixel relevance on one B200 through
the drop-in API (same user code as the reference's examples/vit_torch.py:84-91).  Prints images/s."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lrp-explains-transformers_b200"))

import torch  # noqa: E402
from torchvision.models import vision_transformer  # noqa: E402

from lxt_b200 import ops  # noqa: E402
from lxt_b200.efficient import monkey_patch  # noqa: E402


def main(batch=64, iters=5):
    monkey_patch(vision_transformer)
    torch.manual_seed(0)
    model = vision_transformer.vit_l_16(weights=None)
    torch.nn.init.normal_(model.heads.head.weight, std=0.02)  # torchvision zero-fills the head -> all-zero relevance
    model = model.to(torch.bfloat16).cuda().eval()
    for p in model.parameters():
        p.requires_grad_(False)
    x = torch.randn(batch, 3, 224, 224, generator=torch.Generator().manual_seed(1)).to(torch.bfloat16).cuda()

    def once():
        xi = x.clone().requires_grad_()
        y = model(xi)
        y[torch.arange(batch), y.argmax(-1)].sum().backward()
        return (xi * xi.grad).float().sum(1)

    for _ in range(2):
        heat = once()
    torch.cuda.synchronize()
    n0, t0 = ops.launch_count(), time.perf_counter()
    for _ in range(iters):
        heat = once()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / iters
    print(f"ViT-L/16 bf16, batch {batch}: {dt * 1e3:.1f} ms / batch = {batch / dt:.1f} images/s "
          f"({(ops.launch_count() - n0) // iters} B200 kernel launches per batch); heat-map {tuple(heat.shape)}, "
          f"finite={bool(torch.isfinite(heat).all())}, |heat|={float(heat.norm()):.3e}")


if __name__ == "__main__":
    main()
