"""dcpt_b200.augment on the device: the batched crop + flip / transpose gather gives the CPU result bit for bit (the CPU path is
pinned to the reference's own functions in tests/test_augment_cpu.py)."""
import random

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_crop_augment_on_device_equals_cpu():
    from dcpt_b200.augment import crop_augment_batch, draw_params
    g = torch.Generator().manual_seed(2)
    gt, lq = torch.rand(16, 3, 512, 384, generator=g), torch.rand(16, 3, 256, 192, generator=g)
    params = draw_params(16, 256, 192, 128, random.Random(3))
    a_c, b_c, _ = crop_augment_batch(gt, lq, 256, scale=2, params=params)
    a_g, b_g, _ = crop_augment_batch(gt.cuda(), lq.cuda(), 256, scale=2, params=params)
    assert a_g.is_cuda and tuple(a_g.shape) == (16, 3, 256, 256) and tuple(b_g.shape) == (16, 3, 128, 128)
    assert torch.equal(a_g.cpu(), a_c) and torch.equal(b_g.cpu(), b_c)
