"""BASELINE configs[3] result line: torchvision ViT-L/16 pixel relevance through the drop-in API
(`lxt_b200.efficient.monkey_patch(vision_transformer)`, user code of the reference's examples/vit_torch.py:84-91), bf16, random init,
224 x 224, with and without the Gamma composite (examples/vit_torch.py:59-65).  Prints one JSON line per variant; CUDA events, host
images -> host heat-maps inside the timed region.

    python tools/bench_vit.py [--batch 64] [--steps 5] > profiles/r02_final/bench_vit_l16.jsonl
"""
import argparse
import json
import os
import sys
import warnings

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lrp-explains-transformers_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    from torchvision.models import vision_transformer
    from lxt_b200 import ops
    from lxt_b200.efficient import monkey_patch
    from lxt_b200.efficient.zennit_rules import Gamma, LayerMapComposite
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        monkey_patch(vision_transformer)
    torch.manual_seed(0)
    model = vision_transformer.vit_l_16(weights=None)
    nn.init.normal_(model.heads.head.weight, std=0.02)      # torchvision zero-initialises the head: every logit would be 0
    for p in model.parameters():
        p.requires_grad_(False)
    model = model.cuda().to(torch.bfloat16).eval()
    B = args.batch
    x_host = torch.randn(B, 3, 224, 224, generator=torch.Generator().manual_seed(1)).to(torch.bfloat16).pin_memory()
    heat_host = torch.empty(B, 224, 224, dtype=torch.float32, pin_memory=True)

    def step():
        x = x_host.cuda(non_blocking=True).requires_grad_()
        y = model(x)
        y.max(-1).values.sum().backward()
        heat_host.copy_((x * x.grad).float().sum(1), non_blocking=True)

    for variant in ("plain gradient on Linear / Conv (cp_LRP map only)", "Gamma(conv 100, linear 0.25) composite"):
        comp = None
        if variant.startswith("Gamma"):
            comp = LayerMapComposite([(nn.Conv2d, Gamma(100.0)), (nn.Linear, Gamma(0.25))])
            comp.register(model)
        for _ in range(args.warmup):
            step()
        torch.cuda.synchronize()
        n0 = ops.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        print(json.dumps({"metric": "images/s, ViT-L/16 pixel relevance (224x224, bf16, random init)", "variant": variant, "value": B / ms * 1e3,
                          "unit": "images/s", "batch": B, "ms_per_step": ms, "steps": args.steps,
                          "api": "lxt_b200.efficient.monkey_patch(torchvision.models.vision_transformer)",
                          "launches_per_step": (ops.launch_count() - n0) / args.steps, "finite": bool(torch.isfinite(heat_host).all()),
                          "h2d_bytes_per_step": x_host.numel() * 2, "d2h_bytes_per_step": heat_host.numel() * 4}), flush=True)
        if comp is not None:
            comp.remove()


if __name__ == "__main__":
    main()
