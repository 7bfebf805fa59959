import sys, dataclasses
sys.path.insert(0, "lrp-explains-transformers_b200"); sys.path.insert(0, "tests")
import torch
from helpers import rel_l2
from lxt_b200.engine import LLAMA3_8B, LlamaAttnLRPEngine
dims = dataclasses.replace(LLAMA3_8B, L=4)
eng = LlamaAttnLRPEngine.random_init(dims, device="cuda", seed=0, micro_batch=4)
ids = torch.randint(0, dims.V, (4, 2048), generator=torch.Generator().manual_seed(1)).cuda()
base = eng.attribute_device(ids).clone()
print("device repeats", ["%.2e" % rel_l2(eng.attribute_device(ids), base) for _ in range(8)])
hp = ids.cpu().pin_memory()
print("host api      ", ["%.2e" % rel_l2(eng.attribute(hp), base.cpu()) for _ in range(8)])
r0 = eng.attribute_device(ids[:1]).clone()
print("single prompt ", ["%.2e" % rel_l2(eng.attribute_device(ids[:1]), r0) for _ in range(4)], "vs batch row", "%.2e" % rel_l2(r0, base[:1]))
