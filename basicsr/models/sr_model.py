"""``SRModel``: the fine-tune / test step of the reference (basicsr/models/sr_model.py:27-185, 244-361) on the B200 engines.

``optimize_parameters`` keeps the reference's order - forward, pixel loss, backward, clip, optimizer step, loss log, EMA - and
its option keys (``network_g``, ``path.pretrain_network_g / param_key_g / strict_load_g``, ``train.ema_decay / pixel_opt /
optim_g / scheduler``, top-level ``grad_clip``).  On CUDA the last four stages are ONE fused multi-tensor launch sequence
(``FusedAdam.step(grad_clip=, ema_params=, ema_decay=)``): the clip factor, the Adam update and the EMA blend read each
gradient once.  ``test`` / ``pre_test`` / ``post_test`` / ``test_tile`` are the reference's inference path; tiles of equal shape
go through the network as one batch (dcpt_b200/tiling.py).  LDL and perceptual losses (sr_model.py:142-161) need a VGG /
artifact-map pipeline that is not on this path: asking for them raises."""
from collections import OrderedDict

import torch

from basicsr.archs import build_network
from basicsr.losses import build_loss
from basicsr.utils import get_root_logger
from basicsr.utils.registry import MODEL_REGISTRY

from .base_model import BaseModel


@MODEL_REGISTRY.register()
class SRModel(BaseModel):
    def __init__(self, opt):
        super().__init__(opt)
        self.net_g = self.model_to_device(build_network(opt["network_g"]))
        self.grad_clip = opt.get("grad_clip", 0)
        self._load_pretrained(self.net_g, "g")
        if self.is_train:
            self.init_training_settings()

    def _load_pretrained(self, net, tag, param_key=None):
        paths = self.opt["path"]
        src = paths.get(f"pretrain_network_{tag}", None)
        if src is not None:
            self.load_network(net, src, paths.get(f"strict_load_{tag}", True), param_key or paths.get(f"param_key_{tag}", "params"),
                              self.opt.get("remove_norm", False))
        return src is not None

    def init_training_settings(self):
        self.net_g.train()
        train_opt = self.opt["train"]
        self.ema_decay = train_opt.get("ema_decay", 0)
        if self.ema_decay > 0:
            get_root_logger().info(f"Use Exponential Moving Average with decay: {self.ema_decay}")
            # the EMA copy is never wrapped for data parallel: it is only read by test() and save()
            self.net_g_ema = build_network(self.opt["network_g"]).to(self.device)
            if not self._load_pretrained(self.net_g_ema, "g", param_key="params_ema"):
                self.model_ema(0)
            self.net_g_ema.eval()
        for unsupported in ("ldl_opt", "perceptual_opt"):
            if train_opt.get(unsupported):
                raise NotImplementedError(f"train.{unsupported} is outside the B200 hot path (sr_model.py:142-161)")
        self.cri_pix = build_loss(train_opt["pixel_opt"]).to(self.device) if train_opt.get("pixel_opt") else None
        if self.cri_pix is None:
            raise ValueError("Both pixel and perceptual losses are None.")
        self.setup_optimizers()
        self.setup_schedulers()

    def setup_optimizers(self):
        cfg = dict(self.opt["train"]["optim_g"])
        self.optimizer_g = self.get_optimizer(cfg.pop("type"), self._trainable(self.net_g), **cfg)
        self.optimizers.append(self.optimizer_g)

    def feed_data(self, data):
        self.lq = data["lq"].to(self.device, non_blocking=True)
        if "gt" in data:
            self.gt = data["gt"].to(self.device, non_blocking=True)

    def optimize_parameters(self, current_iter):
        self.net_g.train()
        self.optimizer_g.zero_grad()
        self.output = self.net_g(self.lq)
        l_pix = self.cri_pix(self.output, self.gt)
        loss_dict = OrderedDict(l_pix=l_pix)
        l_pix.backward()
        ema = self.ema_decay > 0
        if self.fused_step():
            # clip_grad_norm_ + Adam(W) + EMA in the fused kernels; trainable parameters only, in named_parameters() order
            ema_params = None
            if ema:
                live = {n for n, p in self.get_bare_model(self.net_g).named_parameters() if p.requires_grad}
                ema_params = [p for n, p in self.net_g_ema.named_parameters() if n in live]
            self.optimizer_g.step(grad_clip=self.grad_clip or None, ema_params=ema_params, ema_decay=self.ema_decay if ema else 0.0)
            if ema and len(ema_params) != sum(1 for _ in self.net_g_ema.parameters()):
                self._ema_frozen()
        else:
            if self.grad_clip:
                torch.nn.utils.clip_grad_norm_(self.net_g.parameters(), self.grad_clip)
            self.optimizer_g.step()
            if ema:
                self.model_ema(decay=self.ema_decay)
        self.prepack(self.net_g)
        self.log_dict = self.reduce_loss_dict(loss_dict)

    def _ema_frozen(self):
        """Frozen parameters are not in the optimizer but the reference's EMA loop covers them too (base_model.py:86-95)."""
        src = dict(self.get_bare_model(self.net_g).named_parameters())
        with torch.no_grad():
            for n, p in self.net_g_ema.named_parameters():
                if not src[n].requires_grad:
                    p.mul_(self.ema_decay).add_(src[n], alpha=1 - self.ema_decay)

    # ------------------------------------------------------------------ inference (sr_model.py:176-185, 244-361)
    def _test_net(self):
        return self.net_g_ema if hasattr(self, "net_g_ema") else self.net_g

    def test(self):
        net = self._test_net()
        was_training = net.training
        net.eval()
        with torch.no_grad():
            self.output = net(self.lq)
        if was_training and net is self.net_g:
            net.train()

    def pre_test(self):
        from dcpt_b200.tiling import pre_test
        self.scale = self.opt.get("scale", 1)
        ws = self.opt["network_g"].get("window_size")
        if ws:
            self.lq, (self.mod_pad_h, self.mod_pad_w) = pre_test(self.lq, ws)
        else:
            self.mod_pad_h = self.mod_pad_w = 0

    def post_test(self):
        from dcpt_b200.tiling import post_test
        self.output = post_test(self.output, (self.mod_pad_h, self.mod_pad_w), self.scale)

    def test_tile(self):
        """``val.infer_size`` x ``val.infer_size`` tiles with ``val.tile_pad`` context (sr_model.py:273-361), batched."""
        from dcpt_b200.tiling import tile_forward
        net = self._test_net()
        was_training = net.training
        net.eval()
        val = self.opt["val"]
        with torch.no_grad():
            self.output = tile_forward(net, self.lq, val["infer_size"], val.get("tile_pad", 32), self.opt.get("scale", 1))
        if was_training and net is self.net_g:
            net.train()

    def get_current_visuals(self):
        out = OrderedDict(lq=self.lq.detach().cpu(), result=self.output.detach().cpu())
        if hasattr(self, "gt"):
            out["gt"] = self.gt.detach().cpu()
        return out

    def save(self, epoch, current_iter):
        if hasattr(self, "net_g_ema"):
            return self.save_network([self.net_g, self.net_g_ema], "net_g", current_iter, param_key=["params", "params_ema"])
        return self.save_network(self.net_g, "net_g", current_iter)
