#!/usr/bin/env python
"""Diagnostic: accuracy and run-to-run reproducibility of torch's fp32 CPU Linear at the oracle's shapes (K = 3072 is the
video input projection) on this host, against float64."""
import hashlib
import torch
import torch.nn.functional as F

g = torch.Generator().manual_seed(0)
for k in (768, 3072):
    x = torch.randn(8192, k, generator=g)
    x = x / x.norm(dim=-1, keepdim=True)
    w = torch.randn(768, k, generator=g) * 0.02
    h = F.layer_norm(x, (k,))
    y = F.linear(h, w)
    y64 = F.linear(F.layer_norm(x.double(), (k,)), w.double())
    err = ((y.double() - y64).abs().max() / y64.abs().max()).item()
    print("K=%d threads=%d  max err / max |y| = %.2e  md5 %s" % (k, torch.get_num_threads(), err,
                                                               hashlib.md5(y.numpy().tobytes()).hexdigest()[:8]))
print(torch.__config__.parallel_info().split("\n")[0], "| mkldnn", torch.backends.mkldnn.is_available(),
      "| fp32 matmul precision", torch.get_float32_matmul_precision())
