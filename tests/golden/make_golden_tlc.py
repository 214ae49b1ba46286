"""Golden vectors for the TLC variant ``NAFNet`` (Local_Base), produced by running the REAL reference on CPU.
Run in the build container only:  ``python tests/golden/make_golden_tlc.py``"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle._ref_import import import_reference  # noqa: E402
from oracle import nafnet_oracle as O  # noqa: E402


def tlc_state_dict(cfg):
    """Seeded weights with the SCA matrices scaled x4, so that local vs global pooling is visible in the output."""
    sd = O.random_nafnet_state_dict(seed=9, **cfg)
    return {k: (v * 4 if "sca.1.weight" in k else v) for k, v in sd.items()}


def tlc_input(g):
    """Large-scale structure (ramps + a checkerboard of 20x16 blocks) + noise: local means differ from the global mean."""
    yy, xx = torch.meshgrid(torch.arange(64.0), torch.arange(80.0), indexing="ij")
    base = torch.stack([xx / 80, yy / 64, ((xx // 20 + yy // 16) % 2)], 0)[None]
    return (0.7 * base + 0.3 * torch.rand(2, 3, 64, 80, generator=g)).clamp(0, 1)


def main():
    import_reference()
    from basicsr.archs.nafnet_arch import NAFNet
    cfg = dict(width=16, enc_blk_nums=[1, 1], middle_blk_num=1, dec_blk_nums=[1, 1])
    train_size = (1, 3, 32, 32)            # base_size 48: level kernels 48 / 24 / 12
    torch.manual_seed(0)
    net = NAFNet(train_size=train_size, **cfg)
    sd = tlc_state_dict(cfg)
    net.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(10)
    inp = tlc_input(g)                             # level maps 64x80 / 32x40 / 16x20: every level pools locally
    with torch.no_grad():
        out = net(inp)
        small = torch.rand(1, 3, 32, 48, generator=g)   # level 0: 32 < 48 rows but 48 >= 48 cols -> global mean at all levels
        out_small = net(small)
    np.savez_compressed(os.path.join(HERE, "nafnet_tlc_w16.npz"), inp=inp.numpy(), out=out.numpy(), small=small.numpy(),
                        out_small=out_small.numpy(), train_size=np.asarray(train_size), seed=9, cfg_width=16, cfg_enc=[1, 1],
                        cfg_mid=1, cfg_dec=[1, 1])
    print("wrote nafnet_tlc_w16.npz")


if __name__ == "__main__":
    main()
