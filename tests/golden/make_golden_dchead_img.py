"""Golden vectors for PromptIR_DC (the classifier head WITH the image embedding, degrad_classify_arch.py:480-556) from the REAL
reference on CPU.  Run in the build container only: ``python tests/golden/make_golden_dchead_img.py``"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle._ref_import import import_reference  # noqa: E402
from oracle import dchead_oracle as D  # noqa: E402


def main():
    import warnings
    warnings.filterwarnings("ignore")
    import_reference()
    from basicsr.archs.degrad_classify_arch import PromptIR_DC
    dims = [16, 32, 32]
    head = PromptIR_DC(feature_dims=dims, num_res_blocks=2, num_classes=5)
    sd = D.random_dchead_state_dict(dims, 2, 5, seed=11, img_embed=True)
    assert [(k, tuple(p.shape)) for k, p in head.named_parameters()] == [(k, tuple(v.shape)) for k, v in sd.items()]   # order + shapes
    head.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(12)
    lq = torch.rand(2, 3, 48, 64, generator=g)
    feats = [torch.randn(2, c, 24 >> i, 32 >> i, generator=g).requires_grad_(True) for i, c in enumerate(dims)]   # features[0] at H/2
    logits = head(lq, list(feats))
    labels = torch.tensor([2, 4])
    loss = torch.nn.functional.cross_entropy(logits, labels)
    loss.backward()
    arrays = {"lq": lq, "logits": logits, "labels": labels, "loss": loss, "dims": dims, "seed": 11}
    for i, f in enumerate(feats):
        arrays[f"feat{i}"] = f
        arrays[f"dfeat{i}"] = f.grad
    for k, p in head.named_parameters():
        arrays["p." + k] = p
        arrays["g." + k] = p.grad
    path = os.path.join(HERE, "dchead_img.npz")
    np.savez_compressed(path, **{k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in arrays.items()})
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
