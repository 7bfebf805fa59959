"""CPU: the NF4 oracle (oracle/nf4_oracle.py) — properties of the published format it restates.  bitsandbytes is absent from this
image and from /root/reference (parity unpinned against bnb); these are the size-independent checks the format offers."""
import numpy as np
import torch

from oracle import nf4_oracle as Q


def test_code_book_is_the_16_normal_quantiles():
    c = Q.NF4_CODE
    assert len(c) == 16 and c[0] == -1.0 and c[7] == 0.0 and c[15] == 1.0
    assert np.all(np.diff(c) > 0)
    assert abs(c[8] - 0.07958029955625534) < 1e-9 and abs(c[1] + 0.6961928009986877) < 1e-9


def test_quantise_dequantise_round_trip_and_idempotence():
    g = torch.Generator().manual_seed(0)
    w = (torch.randn(96, 256, generator=g) * 0.02).to(torch.bfloat16)
    packed, absmax = Q.quantize_nf4(w)
    assert packed.dtype == torch.uint8 and packed.numel() == w.numel() // 2 and absmax.numel() == w.numel() // 64
    wd = Q.dequantize_nf4(packed, absmax, w.shape, dtype=torch.float32)
    # every value lands on code * absmax of its block; the block maximum is reproduced exactly (codes +-1)
    blk = w.float().reshape(-1, 64)
    assert torch.equal(wd.reshape(-1, 64).abs().amax(1), blk.abs().amax(1))
    # worst-case error: half the largest code gap (0.3039 / 2) times absmax
    err = (wd - w.float()).reshape(-1, 64).abs().amax(1)
    assert bool((err <= 0.152 * absmax + 1e-12).all())
    rel = float((wd - w.float()).norm() / w.float().norm())
    assert 0.05 < rel < 0.12          # ~0.09 for Gaussian weights: the known NF4 noise level
    # idempotence: quantising the de-quantised weight gives the same codes (fp32 values: no second bf16 rounding)
    p2, a2 = Q.quantize_nf4(wd.to(torch.bfloat16))
    wd2 = Q.dequantize_nf4(p2, a2, w.shape, dtype=torch.float32)
    assert float((wd2 - wd).abs().max()) <= float(absmax.max()) * 0.01


def test_packing_puts_the_first_value_in_the_high_nibble():
    w = torch.zeros(64)
    w[0], w[1] = 1.0, -1.0           # codes 15 and 0
    packed, absmax = Q.quantize_nf4(w)
    assert int(packed[0]) == (15 << 4) | 0 and float(absmax[0]) == 1.0
    assert int(packed[1]) == (7 << 4) | 7      # zeros -> code 7
