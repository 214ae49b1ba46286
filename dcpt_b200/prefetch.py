"""Device-side input staging for the training loop (reference: basicsr/data/prefetch_dataloader.py:83-125 ``CUDAPrefetcher``,
selected by ``prefetch_mode: cuda`` in basicsr/train.py): while step i runs, the host->device copy of batch i+1 is already in
flight on a second stream, so the PCIe transfer (25 MB per step for 16 x 3 x 256 x 256 lq + gt) leaves the critical path.

Same interface and semantics as the reference class (``next()`` returns the staged batch dict or None at the end of the
epoch and starts staging the following one; ``reset()`` restarts the epoch).  Differences: the staged tensors are tied to the
consumer stream with ``record_stream`` (the reference relies on the allocator not reusing them early), dict values that are
already on the device pass through, and the batches should sit in pinned memory for the copy to be asynchronous (the
reference's loaders set ``pin_memory: true``)."""
import torch


class CUDAPrefetcher:
    def __init__(self, loader, opt=None, device=None):
        self.ori_loader = loader
        self.loader = iter(loader)
        self.opt = opt
        if device is None:
            device = torch.device("cuda" if (opt is None or opt.get("num_gpu", 1) != 0) else "cpu")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            from .lib import DcptError
            raise DcptError("CUDAPrefetcher stages batches on a CUDA device (prefetch_mode: cuda); there is no CPU path")
        self.stream = torch.cuda.Stream(device=self.device)
        self.batch = None
        self.preload()

    def preload(self):
        try:
            batch = next(self.loader)
        except StopIteration:
            self.batch = None
            return None
        with torch.cuda.stream(self.stream):
            self.batch = {k: (v.to(device=self.device, non_blocking=True) if torch.is_tensor(v) else v) for k, v in batch.items()}
        return None

    def next(self):
        cur = torch.cuda.current_stream(self.device)
        cur.wait_stream(self.stream)
        batch = self.batch
        if batch is not None:
            for v in batch.values():
                if torch.is_tensor(v) and v.is_cuda:
                    v.record_stream(cur)
        self.preload()
        return batch

    def reset(self):
        self.loader = iter(self.ori_loader)
        self.preload()
