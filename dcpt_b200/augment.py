"""Device-side crop + flip / transpose augmentation for a BATCH of training pairs (first piece).

The reference augments one sample at a time on the CPU inside ``PairedImageDataset.__getitem__``
(basicsr/data/paired_image_dataset.py) with ``paired_random_crop`` (basicsr/data/transforms.py:48-129: one random top / left per
sample, the GT window is the LQ window times ``scale``) followed by ``augment`` (:132-196 - in this fork: a random code in {0, 1, 2}
= horizontal flip / vertical flip / none, and a transpose with probability 3/4; its ``hflip`` / ``rotation`` arguments are ignored).
Here the whole batch is cut and augmented by ONE gather per tensor on the device the images already live on: per sample an index map
says which source pixel every output pixel reads, so crop, flip and transpose cost one pass over the patch instead of three over the
image.  The random draws stay on the host (``draw_params`` uses Python's ``random`` in the reference's order: top, left, flip code,
rotation code per sample); only indices travel.  Pure index arithmetic: bit-exact against the reference functions on the same draws
(tests/test_augment_cpu.py)."""
import random

import torch


def draw_params(batch, h_lq, w_lq, lq_patch_size, rng=random):
    """Per-sample (top, left, flip_code, rot_code) drawn like the reference does for one sample after the other:
    ``random.randint(0, h_lq - p)``, ``random.randint(0, w_lq - p)`` (transforms.py:98-99), then ``random.randint(0, 2)``,
    ``random.randint(0, 3)`` (:154-155)."""
    if h_lq < lq_patch_size or w_lq < lq_patch_size:
        raise ValueError(f"LQ ({h_lq}, {w_lq}) is smaller than patch size ({lq_patch_size}, {lq_patch_size}).")
    out = []
    for _ in range(batch):
        top, left = rng.randint(0, h_lq - lq_patch_size), rng.randint(0, w_lq - lq_patch_size)
        out.append((top, left, rng.randint(0, 2), rng.randint(0, 3)))
    return out


def _index_maps(params, patch, scale, device):
    """[B, patch, patch] source row / column of every output pixel: crop window, then flip, then transpose (the reference's order:
    ``cv2.flip`` in place, then ``img.transpose(1, 0, 2)``)."""
    B = len(params)
    p = torch.tensor(params, dtype=torch.long, device=device)               # [B, 4]
    top, left, flip, rot = (p[:, 0] * scale).view(B, 1, 1), (p[:, 1] * scale).view(B, 1, 1), p[:, 2].view(B, 1, 1), p[:, 3].view(B, 1, 1)
    ar = torch.arange(patch, device=device)
    oy, ox = ar.view(1, patch, 1).expand(B, patch, patch), ar.view(1, 1, patch).expand(B, patch, patch)
    tr = rot != 0
    fy, fx = torch.where(tr, ox, oy), torch.where(tr, oy, ox)               # undo the transpose: out[y, x] = flipped[x, y]
    cy = torch.where(flip == 1, patch - 1 - fy, fy)                         # undo the vertical flip
    cx = torch.where(flip == 0, patch - 1 - fx, fx)                         # undo the horizontal flip
    return top + cy, left + cx


def crop_augment_batch(gt, lq, gt_patch_size, scale=1, params=None, rng=random):
    """gt [B, C, H, W], lq [B, C, H / scale, W / scale] (same device) -> (gt patches [B, C, P, P], lq patches [B, C, P / scale, P / scale],
    params).  Equivalent to ``paired_random_crop`` + ``augment`` of the reference applied per sample with the draws in ``params``."""
    if gt.dim() != 4 or lq.dim() != 4 or gt.shape[0] != lq.shape[0]:
        raise ValueError("crop_augment_batch expects batched [B, C, H, W] tensors of equal batch size")
    B, _, h_gt, w_gt = gt.shape
    h_lq, w_lq = lq.shape[-2:]
    if h_gt != h_lq * scale or w_gt != w_lq * scale:
        raise ValueError(f"Scale mismatches. GT ({h_gt}, {w_gt}) is not {scale}x multiplication of LQ ({h_lq}, {w_lq}).")
    lq_patch = gt_patch_size // scale
    if params is None:
        params = draw_params(B, h_lq, w_lq, lq_patch, rng)
    bi = torch.arange(B, device=gt.device).view(B, 1, 1)

    def gather(x, patch, s):
        yy, xx = _index_maps(params, patch, s, x.device)
        return x[bi, :, yy, xx].permute(0, 3, 1, 2).contiguous()            # advanced indexing puts the channel axis last

    return gather(gt, gt_patch_size, scale), gather(lq, lq_patch, 1), params
