"""``DCPTModel``: the degradation-classification pretrain step (reference:
basicsr/models/degradation_classification_pretrain_model.py:16-169) on the B200 engines.

One iteration, in the reference's order (:133-169): restoration pass ``net_g(gt, hook=False)`` -> pixel loss against ``gt``;
hooked pass ``net_g(lq, hook=True)`` (returns None, the decoder-stage outputs arrive through forward hooks on the modules whose
name contains ``hook_names`` and has exactly one dot, :64-67) -> ``net_dc(lq, hook_outputs[::-1])`` -> classification loss;
ONE backward of the sum; both optimizers step.  The two net_g passes are two engine backward nodes accumulating into the same
``.grad`` tensors.  Under data parallel the per-node all-reduce of the wrapper is therefore switched off for the backward and
every network's accumulated gradient buffer is exchanged once afterwards (``dcpt_b200.dist.exchange_accumulated_grads_``):
one 272 MB + one 197 MB all-reduce per step instead of three."""
from collections import OrderedDict
from contextlib import ExitStack

from basicsr.archs import build_network
from basicsr.losses import build_loss
from basicsr.utils.registry import MODEL_REGISTRY

from .base_model import BaseModel


@MODEL_REGISTRY.register()
class DCPTModel(BaseModel):
    def __init__(self, opt):
        super().__init__(opt)
        self.net_g = self.model_to_device(build_network(opt["network_g"]))
        self.net_dc = self.model_to_device(build_network(opt["network_dc"]))
        paths = opt["path"]
        for tag, net in (("g", self.net_g), ("dc", self.net_dc)):
            src = paths.get(f"pretrain_network_{tag}", None)
            if src is not None:
                self.load_network(net, src, paths.get(f"strict_load_{tag}", True), paths.get(f"param_key_{tag}", "params"),
                                  opt.get("remove_norm", False))
        if self.is_train:
            self.init_training_settings()

    def init_training_settings(self):
        self.net_g.train()
        self.net_dc.train()
        train_opt = self.opt["train"]
        self.hook_outputs, self.hooks = [], []
        pattern = self.opt.get("hook_names", None)
        for name, module in self.net_g.named_modules():
            if pattern in name and name.count(".") == 1:      # one-dot rule (a DP wrapper's "module." prefix supplies the dot)
                self.hooks.append(module.register_forward_hook(self.hook_forward_fn))
        self.cri_classify = build_loss(train_opt["classify_opt"]).to(self.device) if train_opt.get("classify_opt") else None
        self.cri_pixel = build_loss(train_opt["pixel_opt"]).to(self.device) if train_opt.get("pixel_opt") else None
        if self.cri_classify is None:
            raise ValueError("Classify loss is None.")
        self.setup_optimizers()
        self.setup_schedulers()

    def hook_forward_fn(self, module, input, output):  # noqa: A002
        self.hook_outputs.append(output[-1] if isinstance(output, tuple) else output)

    def setup_optimizers(self):
        for tag, net in (("g", self.net_g), ("dc", self.net_dc)):
            cfg = dict(self.opt["train"][f"optim_{tag}"])
            optimizer = self.get_optimizer(cfg.pop("type"), self._trainable(net), **cfg)
            setattr(self, f"optimizer_{tag}", optimizer)
            self.optimizers.append(optimizer)

    def feed_data(self, data):
        self.lq = data["lq"].to(self.device, non_blocking=True)
        self.dataset_idx = data["dataset_idx"].to(self.device, non_blocking=True)
        if "gt" in data:
            self.gt = data["gt"].to(self.device, non_blocking=True)

    def optimize_parameters(self, current_iter):
        loss_dict = OrderedDict()
        # restoration of the clean image
        self.net_g.train()
        self.net_dc.eval()
        self.optimizer_g.zero_grad()
        pix_output = self.net_g(self.gt, hook=False)
        self.hook_outputs = []
        l_total = 0
        if self.cri_pixel:
            loss_dict["l_pix"] = self.cri_pixel(pix_output, self.gt)
            l_total = l_total + loss_dict["l_pix"]
        # degradation classification from the decoder features of the degraded image
        self.net_dc.train()
        self.optimizer_dc.zero_grad()
        self.net_g(self.lq, hook=True)
        cls_output = self.net_dc(self.lq, self.hook_outputs[::-1])
        loss_dict["l_classify"] = self.cri_classify(cls_output, self.dataset_idx)
        l_total = l_total + loss_dict["l_classify"]
        wrapped = [n for n in (self.net_g, self.net_dc) if hasattr(n, "no_sync")]
        with ExitStack() as stack:
            for n in wrapped:
                stack.enter_context(n.no_sync())
            l_total.backward()
        if wrapped:
            from dcpt_b200.dist import exchange_accumulated_grads_
            exchange_accumulated_grads_(wrapped)
        self.optimizer_g.step()
        self.optimizer_dc.step()
        self.prepack(self.net_g, self.net_dc)
        self.hook_outputs = []
        self.log_dict = self.reduce_loss_dict(loss_dict)

    def save(self, epoch, current_iter):
        return [self.save_network(self.net_g, "net_g", current_iter), self.save_network(self.net_dc, "net_dc", current_iter)]
