"""PromptIR on the B200 hot path - drop-in for the reference's ``basicsr/archs/promptir_arch.py`` (inference).

Same class name, ctor kwargs, sub-module / parameter names and ``state_dict`` shapes as the reference (promptir_arch.py:238-263
PromptGenBlock, :267-462 PromptIR), so ``options/all_in_one/test/test_PromptIR_5d.yml``'s ``network_g: {type: PromptIR,
window_size: 8}`` and its ``params_ema`` checkpoint load with ``strict=True``.  The modules are parameter containers;
``PromptIR.forward`` hands the whole network to the sm_100a kernels through the C ABI (``dcpt_promptir_fwd``).  The transformer
blocks are the Restormer mirror's containers: PromptIR's differ only in the softmax on the attention map (:140), which is a flag
of the engine, not a parameter."""
import torch
import torch.nn as nn

from basicsr.utils.registry import ARCH_REGISTRY
from dcpt_b200.lib import DcptError
from dcpt_b200.promptir import PromptIREngine

from .restormer_arch import Downsample, OverlapPatchEmbed, TransformerBlock, Upsample


class PromptGenBlock(nn.Module):
    """promptir_arch.py:238-250: ``prompt_len`` learned prompt components, a linear layer that mixes them from the globally
    pooled features, a 3x3 conv on the resized mixture."""

    def __init__(self, prompt_dim=128, prompt_len=5, prompt_size=96, lin_dim=192):
        super().__init__()
        self.prompt_param = nn.Parameter(torch.rand(1, prompt_len, prompt_dim, prompt_size, prompt_size), requires_grad=True)
        self.linear_layer = nn.Linear(lin_dim, prompt_len)
        self.conv3x3 = nn.Conv2d(prompt_dim, prompt_dim, kernel_size=3, stride=1, padding=1, bias=False)


@ARCH_REGISTRY.register()
class PromptIR(nn.Module):
    """promptir_arch.py:267-518.  ``window_size`` is accepted and ignored by the arch, as in the reference."""

    def __init__(self, inp_channels=3, out_channels=3, dim=48, num_blocks=[4, 6, 6, 8], num_refinement_blocks=4, heads=[1, 2, 4, 8],
                 ffn_expansion_factor=2.66, bias=False, LayerNorm_type="WithBias", decoder=True, window_size=8):
        super().__init__()
        if not decoder:
            raise DcptError("PromptIR(decoder=False) is not built: without the prompt path the reference's channel counts do not "
                            "assemble either (reduce_chan_level3 expects 2*dim+192 inputs, promptir_arch.py:366-368)")
        if bias:
            raise DcptError("PromptIR(bias=True) is not built: unlike the fork's Restormer, PromptIR's attention / feed-forward convs "
                            "would carry biases then (promptir_arch.py:82-98, 114-127); no shipped config sets it")
        a = (ffn_expansion_factor, bias, LayerNorm_type)
        stage = lambda d, h, n: nn.Sequential(*[TransformerBlock(d, h, *a) for _ in range(n)])   # noqa: E731
        self.patch_embed = OverlapPatchEmbed(inp_channels, dim)
        self.decoder = decoder
        self.prompt1 = PromptGenBlock(prompt_dim=64, prompt_len=5, prompt_size=64, lin_dim=96)
        self.prompt2 = PromptGenBlock(prompt_dim=128, prompt_len=5, prompt_size=32, lin_dim=192)
        self.prompt3 = PromptGenBlock(prompt_dim=320, prompt_len=5, prompt_size=16, lin_dim=384)
        self.encoder_level1 = stage(dim, heads[0], num_blocks[0])
        self.down1_2 = Downsample(dim)
        self.encoder_level2 = stage(dim * 2, heads[1], num_blocks[1])
        self.down2_3 = Downsample(dim * 2)
        self.encoder_level3 = stage(dim * 4, heads[2], num_blocks[2])
        self.down3_4 = Downsample(dim * 4)
        self.latent = stage(dim * 8, heads[3], num_blocks[3])
        self.up4_3 = Upsample(dim * 4)
        self.reduce_chan_level3 = nn.Conv2d(dim * 2 + 192, dim * 4, kernel_size=1, bias=bias)
        self.noise_level3 = TransformerBlock(dim * 4 + 512, heads[2], *a)
        self.reduce_noise_level3 = nn.Conv2d(dim * 4 + 512, dim * 4, kernel_size=1, bias=bias)
        self.decoder_level3 = stage(dim * 4, heads[2], num_blocks[2])
        self.up3_2 = Upsample(dim * 4)
        self.reduce_chan_level2 = nn.Conv2d(dim * 4, dim * 2, kernel_size=1, bias=bias)
        self.noise_level2 = TransformerBlock(dim * 2 + 224, heads[2], *a)
        self.reduce_noise_level2 = nn.Conv2d(dim * 2 + 224, dim * 4, kernel_size=1, bias=bias)
        self.decoder_level2 = stage(dim * 2, heads[1], num_blocks[1])
        self.up2_1 = Upsample(dim * 2)
        self.noise_level1 = TransformerBlock(dim * 2 + 64, heads[2], *a)
        self.reduce_noise_level1 = nn.Conv2d(dim * 2 + 64, dim * 2, kernel_size=1, bias=bias)
        self.decoder_level1 = stage(dim * 2, heads[0], num_blocks[0])
        self.refinement = stage(dim * 2, heads[0], num_refinement_blocks)
        self.output = nn.Conv2d(dim * 2, out_channels, kernel_size=3, stride=1, padding=1, bias=bias)
        self._cfg = dict(inp_channels=inp_channels, out_channels=out_channels, dim=dim, num_blocks=tuple(num_blocks),
                         num_refinement_blocks=num_refinement_blocks, heads=tuple(heads), ffn_expansion_factor=ffn_expansion_factor,
                         bias=bias, ln_with_bias=LayerNorm_type != "BiasFree")
        self._engine = None

    def engine(self):
        if self._engine is None:
            self._engine = PromptIREngine(**self._cfg)
        return self._engine

    def forward(self, inp_img, hook=False):
        params = list(self.parameters())
        if torch.is_grad_enabled() and (inp_img.requires_grad or any(p.requires_grad for p in params)):
            raise DcptError("PromptIR on the B200 path is inference-only: call it under torch.no_grad() (SRModel.test does)")
        if hook:
            raise DcptError("PromptIR(hook=True) belongs to DCPT pretraining of PromptIR, which is not built")
        return self.engine().forward([p.detach() for p in params], inp_img)
