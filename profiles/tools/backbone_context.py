"""Context only: how long the (out-of-scope, stock PyTorch) backbone takes next to the hot path.
torchvision stand-ins for the BASELINE backbones, random init, bf16 autocast, channels_last, batch 512."""
import json
import sys

import torch
import torchvision

dev = torch.device("cuda:0")
out = {}
for name, ctor in (("resnet50", torchvision.models.resnet50), ("resnet18", torchvision.models.resnet18),
                   ("efficientnet_b0", torchvision.models.efficientnet_b0), ("convnext_tiny", torchvision.models.convnext_tiny),
                   ("vit_b_16", torchvision.models.vit_b_16)):
    m = ctor(weights=None).to(dev).eval().to(memory_format=torch.channels_last)
    x = torch.randn(512, 3, 224, 224, device=dev).to(memory_format=torch.channels_last)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        for _ in range(3):
            m(x)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            m(x)
        b.record()
        torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    out[name] = {"fwd_ms_per_512": ms, "crops_per_s": 512 / ms * 1e3}
    del m, x
    torch.cuda.empty_cache()
print(json.dumps(out))
