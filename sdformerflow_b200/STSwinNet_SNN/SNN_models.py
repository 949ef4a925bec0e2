"""Multi-resolution spiking U-Net plumbing around the Swin encoder (host-side mirror of reference
models/STSwinNet_SNN/SNN_models.py:12-216).  Thin glue: residual blocks, transposed-conv decoders
and prediction layers built from Spiking_modules (their convolutions run on the spike GEMM engine, ops.spike_conv_gemm / spike_deconv)."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from .Spiking_modules import *  # noqa: F401,F403
from .Spiking_modules import (SpikingConvEncoderLayer, SEWResBlock, SpikingTransposeDecoderLayer, SpikingPredLayer,
                              MS_SpikingConvEncoderLayer, MS_ResBlock, MS_SpikingTransposeDecoderLayer,
                              MS_SpikingPredLayer)


def skip_concat(x1, x2, dim=1):
    """Zero-pad x1 to x2's spatial size and concatenate (reference models/model_util.py:13-18)."""
    dY, dX = x2.size(-2) - x1.size(-2), x2.size(-1) - x1.size(-1)
    return torch.cat([F.pad(x1, (dX // 2, dX - dX // 2, dY // 2, dY - dY // 2)), x2], dim=dim)


def skip_concat_cl(x1, x2, align=1):
    """skip_concat on channels-last (B, T, H, W, C) tensors: pad x1's H, W to x2's, concatenate channels.
    align > 1: zero channels are appended up to a multiple of `align` in the same pass (the consumer pads its weight with
    zero input slices: Spiking_modules.padded_in_weight) — 770 / 386 / 194-channel decoder inputs otherwise cost cuDNN a
    padding copy of the whole tensor per convolution call."""
    dY, dX = x2.size(2) - x1.size(2), x2.size(3) - x1.size(3)
    if dY or dX:
        x1 = F.pad(x1, (0, 0, dX // 2, dX - dX // 2, dY // 2, dY - dY // 2))
    parts = [x1, x2]
    extra = -(x1.size(-1) + x2.size(-1)) % align
    if extra:
        parts.append(x2.new_zeros(*x2.shape[:-1], extra))
    return torch.cat(parts, dim=-1)


def skip_sum(x1, x2, dim=None):
    dY, dX = x2.size(-2) - x1.size(-2), x2.size(-1) - x1.size(-1)
    return F.pad(x1, (dX // 2, dX - dX // 2, dY // 2, dY - dY // 2)) + x2


class SpikingMultiResUNet(nn.Module):
    """Builder of the residual / decoder / prediction stacks (reference SNN_models.py:12-153).
    The conv-encoder variant of the reference (build_encoders + its forward) is not part of the
    Swin hot path; subclasses provide the encoder."""
    ff_type = SpikingConvEncoderLayer
    res_type = SEWResBlock
    upsample_type = None          # upsample-conv decoder (SpikingDecoderLayer) not built: shipped configs use transposed conv
    transpose_type = SpikingTransposeDecoderLayer
    pred_type = SpikingPredLayer
    input_sfn = True
    w_scale_pred = 0.01
    upsample_4 = False

    def __init__(self, base_num_channels, num_encoders, num_residual_blocks, num_output_channels, skip_type, norm,
                 use_upsample_conv, num_bins, recurrent_block_type=None, kernel_size=5, channel_multiplier=2,
                 activations=("relu", None), final_activation=None, spiking_neuron=None):
        super().__init__()
        self.base_num_channels, self.num_encoders = base_num_channels, num_encoders
        self.num_residual_blocks, self.num_output_channels = num_residual_blocks, num_output_channels
        self.kernel_size, self.skip_type = kernel_size, skip_type
        self.norm = None
        self.recurrent_block_type, self.channel_multiplier = recurrent_block_type, channel_multiplier
        self.ff_act, self.rec_act = activations
        self.final_activation, self.num_bins_all = final_activation, num_bins
        self.spiking_kwargs = {}
        if type(spiking_neuron) is dict:
            self.spiking_kwargs.update(spiking_neuron)
            self.steps = self.spiking_kwargs["num_steps"]
            self.num_ch = num_bins * 2 // self.steps
        self.skip_ftn = {"concat": skip_concat, "sum": skip_sum}[skip_type]
        if use_upsample_conv:
            if self.upsample_type is None:
                raise NotImplementedError("use_upsample_conv=True (bilinear upsample + conv decoder) is not built; the "
                                          "shipped SNN configs set use_upsample_conv: False")
            self.UpsampleLayer = self.upsample_type
        else:
            self.UpsampleLayer = self.transpose_type
        assert self.num_output_channels > 0
        m = self.channel_multiplier
        self.encoder_input_sizes = [int(base_num_channels * pow(m, i)) for i in range(num_encoders)]
        self.encoder_output_sizes = [int(base_num_channels * pow(m, i + 1)) for i in range(num_encoders)]
        self.max_num_channels = self.encoder_output_sizes[-1]

    def build_resblocks(self):
        return nn.ModuleList([self.res_type(self.max_num_channels, self.max_num_channels, connect_function="ADD",
                                            **self.spiking_kwargs) for _ in range(self.num_residual_blocks)])

    def build_multires_prediction_layer(self):
        return nn.ModuleList([self.pred_type(c, self.num_output_channels, 1, **self.spiking_kwargs)
                              for c in reversed(self.encoder_input_sizes)])

    def build_multires_prediction_decoders(self):
        decoders = nn.ModuleList()
        i_max = len(self.encoder_input_sizes) - 1
        pairs = zip(reversed(self.encoder_output_sizes), reversed(self.encoder_input_sizes))
        for i, (cin, cout) in enumerate(pairs):
            sf = 4 if (self.upsample_4 and i == i_max) else 2
            decoders.append(self.UpsampleLayer(2 * cin + (0 if i == 0 else self.num_output_channels), cout,
                                               kernel_size=self.kernel_size, scale=sf, **self.spiking_kwargs))
        return decoders
