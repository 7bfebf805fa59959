"""CPU oracle for the 4-bit NormalFloat (NF4) weight format  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restates the published NF4 definition (QLoRA, Dettmers et al. 2023, the format bitsandbytes' `Linear4bit(quant_type="nf4")`
stores: 16 normal quantiles scaled to [-1, 1], one fp32 absmax per block of 64 values, two codes per byte with the first value
in the high nibble).  The reference only ever *uses* such weights through bitsandbytes (examples/quantized_llama.py:13-19,
lxt/explicit/models/llama.py:91-92); bitsandbytes is absent from this image and from /root/reference, so:
**parity unpinned against bitsandbytes** — the CUDA quantiser / de-quantiser (csrc/quant.cu) are pinned bit-exactly against THIS
restatement, and the engine running on NF4 weights is pinned against the fp32 AttnLRP oracle evaluated on the de-quantised weights.
"""
import numpy as np
import torch

NF4_CODE = np.array([-1.0, -0.6961928009986877, -0.5250730514526367, -0.39491748809814453, -0.28444138169288635,
                     -0.18477343022823334, -0.09105003625154495, 0.0, 0.07958029955625534, 0.16093020141124725,
                     0.24611230194568634, 0.33791524171829224, 0.44070982933044434, 0.5626170039176941, 0.7229568362236023, 1.0],
                    dtype=np.float32)


def quantize_nf4(w: torch.Tensor, blocksize: int = 64):
    """w (any float dtype, values taken as bf16-rounded fp32) -> (packed uint8 [n/2], absmax fp32 [n/blocksize])"""
    x = w.detach().to(torch.bfloat16).float().cpu().numpy().reshape(-1, blocksize)
    absmax = np.abs(x).max(axis=1).astype(np.float32)
    inv = np.where(absmax > 0, np.float32(1.0) / absmax, np.float32(0.0)).astype(np.float32)
    xn = (x * inv[:, None]).astype(np.float32)
    dist = np.abs(xn[:, :, None] - NF4_CODE[None, None, :])
    code = dist.argmin(axis=2).astype(np.uint8)            # first minimum = lowest code index on ties
    code = code.reshape(-1)
    packed = (code[0::2] << 4) | code[1::2]
    return torch.from_numpy(packed.astype(np.uint8)), torch.from_numpy(absmax)


def dequantize_nf4(packed: torch.Tensor, absmax: torch.Tensor, shape, blocksize: int = 64, dtype=torch.bfloat16) -> torch.Tensor:
    p = packed.cpu().numpy()
    code = np.empty(p.size * 2, dtype=np.uint8)
    code[0::2] = p >> 4
    code[1::2] = p & 15
    vals = NF4_CODE[code].reshape(-1, blocksize) * absmax.cpu().numpy().astype(np.float32)[:, None]
    return torch.from_numpy(vals.reshape(shape).astype(np.float32)).to(dtype)
