"""Golden vectors for the PromptIR transformer block (softmax MDTA), produced by running the REAL reference
(``/root/reference/basicsr/archs/promptir_arch.py:108-186``) on CPU.  Run in the build container only:
``python tests/golden/make_golden_promptir.py``"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle._ref_import import import_reference  # noqa: E402


def main():
    import_reference()
    from basicsr.archs.promptir_arch import TransformerBlock
    torch.set_num_threads(4)
    for dim, heads, hw, lnt in ((48, 1, (16, 24), "WithBias"), (96, 2, (8, 16), "WithBias"), (64, 4, (8, 8), "BiasFree")):
        torch.manual_seed(300 + dim)
        blk = TransformerBlock(dim, heads, 2.66, False, lnt)
        g = torch.Generator().manual_seed(400 + dim)
        sd = {}
        for k, p in blk.named_parameters():
            shp = tuple(p.shape)
            if k.endswith("temperature"):
                sd[k] = 0.5 + 4.0 * torch.rand(shp, generator=g)     # sharp enough that the softmax is far from uniform
            elif "norm" in k and k.endswith("weight"):
                sd[k] = 1.0 + 0.1 * torch.randn(shp, generator=g)
            elif k.endswith("bias"):
                sd[k] = 0.1 * torch.randn(shp, generator=g)
            else:
                sd[k] = torch.randn(shp, generator=g) / (shp[1] * shp[2] * shp[3]) ** 0.5
        blk.load_state_dict(sd, strict=True)
        x = torch.randn(2, dim, *hw).requires_grad_(True)
        y = blk(x)
        dy = torch.randn_like(y)
        y.backward(dy)
        arrays = {"x": x, "y": y, "dy": dy, "dx": x.grad, "heads": heads, "seed": 300 + dim, "ln_bias": int(lnt != "BiasFree")}
        for k, p in blk.named_parameters():
            arrays["p." + k] = p
            arrays["g." + k] = p.grad
        path = os.path.join(HERE, f"promptir_block_d{dim}.npz")
        np.savez_compressed(path, **{k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in arrays.items()})
        print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
