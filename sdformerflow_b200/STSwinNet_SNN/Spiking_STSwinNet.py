"""SNN optical-flow networks: Swin spiking encoder + conv decoder (host-side mirror of reference
models/STSwinNet_SNN/Spiking_STSwinNet.py; FlowNet kwargs assembly restated from
models/STSwinNet/STSwinNet.py:323-366).

Drop-in surface: ``Model(config["model"].copy(), config["swin_transformer"].copy())``,
``.init_weights()``, ``model(chunk)`` with chunk (B, num_bins, 2, H, W) ->
{"flow": [(B, 2, H, W)] * num_encoders, "attn": None}; ``state_dict`` keys as SURVEY.md Appendix D.
"""
import torch
from torch import nn

from .. import ops
from ..sj import layer, functional, neuron  # noqa: F401
from .Spiking_swin_transformer3D import Spiking_SwinTransformer3D_v2, MS_Spiking_SwinTransformer3D_v2
from .SNN_models import *  # noqa: F401,F403
from .SNN_models import SpikingMultiResUNet, skip_concat_cl
from .Spiking_modules import (MS_SpikingConvEncoderLayer, MS_ResBlock, MS_SpikingTransposeDecoderLayer,
                              MS_SpikingPredLayer)


class spiking_former_encoder(nn.Module):
    """Swin3D encoder wrapper (reference :8-85).  Output: per-stage features as (T, B, C, H, W) views."""
    swin_type = Spiking_SwinTransformer3D_v2

    def __init__(self, arc_type="swinv2", patch_embed_type="PatchEmbedLocal", img_size=(240, 320), patch_size=(32, 2, 2),
                 in_chans=128, embed_dim=96, depths=[2, 2, 6], num_heads=[3, 6, 12], window_size=[2, 7, 7],
                 pretrained_window_size=[0, 0, 0], mlp_ratio=4.0, patch_norm=False, out_indices=(0, 1, 2),
                 frozen_stages=-1, norm=None, spikformer_norm=None, pol_in_channel=False, **spiking_kwargs):
        super().__init__()
        self.num_blocks = in_chans // patch_size[0]
        self.img_size, self.patch_size, self.in_chans, self.embed_dim = img_size, patch_size, in_chans, embed_dim
        self.depths, self.num_heads, self.patch_norm = depths, num_heads, patch_norm
        self.window_size, self.mlp_ratio, self.out_indices = window_size, mlp_ratio, out_indices
        self.frozen_stages, self.num_encoders = frozen_stages, len(depths)
        self.out_channels = [embed_dim * (2 ** i) for i in range(self.num_encoders)]
        self.spikformer_norm = spikformer_norm
        self.swin3d = self.swin_type(
            arc_type=arc_type, embed_type=patch_embed_type, img_size=img_size, patch_size=patch_size, in_chans=in_chans,
            embed_dim=embed_dim, depths=depths, num_heads=num_heads, window_size=window_size,
            pretrained_window_size=pretrained_window_size, mlp_ratio=mlp_ratio, drop_rate=0.0, attn_drop_rate=0.0,
            drop_path_rate=0.2, norm_layer=spikformer_norm, out_indices=out_indices, frozen_stages=frozen_stages,
            norm=norm, **spiking_kwargs)

    def forward_cl(self, inputs):
        return self.swin3d.features_cl(inputs)                        # (B, D, Hi, Wi, Ci) per stage

    def forward(self, inputs):
        feats = self.swin3d(inputs)                                   # (B, C, D, H, W) views
        return [feats[i].permute(2, 0, 1, 3, 4) for i in range(self.num_encoders)]


class MS_spiking_former_encoder(spiking_former_encoder):
    swin_type = MS_Spiking_SwinTransformer3D_v2


class Spikingformer_MultiResUNet(SpikingMultiResUNet):
    """U-Net with a spiking Swin encoder and transposed-conv spiking decoders, SEW shortcut (reference :88-252)."""
    pol_channel = False
    encoder_block = spiking_former_encoder
    upsample_4 = False

    def __init__(self, unet_kwargs, stt_kwargs):
        unet_kwargs.pop("spiking_feedforward_block_type", None)
        super().__init__(**unet_kwargs)
        self.arc_type, self.patch_embed_type = stt_kwargs["use_arc"][0], stt_kwargs["use_arc"][1]
        self.num_bins_events = unet_kwargs["num_bins"]
        ints = lambda key: [int(i) for i in stt_kwargs[key]]  # noqa: E731
        self.depths, self.num_heads = ints("swin_depths"), ints("swin_num_heads")
        assert len(self.depths) == self.num_encoders and len(self.num_heads) == self.num_encoders
        self.patch_size, self.out_indices = ints("swin_patch_size"), ints("swin_out_indices")
        self.window_size, self.pretrained_window_size = ints("window_size"), ints("pretrained_window_size")
        self.mlp_ratio, self.input_size = stt_kwargs["mlp_ratio"], stt_kwargs["input_size"]
        self.spikformer_norm = stt_kwargs["norm"] if "norm" in stt_kwargs else unet_kwargs["spiking_neuron"]["spike_norm"]
        m = self.channel_multiplier
        self.encoder_output_sizes = [int(self.base_num_channels * pow(m, i)) for i in range(self.num_encoders)]
        self.encoder_input_sizes = [self.base_num_channels] + self.encoder_output_sizes[:-1]
        self.max_num_channels = self.encoder_output_sizes[-1]
        self.resblocks = self.build_resblocks()
        self.decoders = self.build_multires_prediction_decoders()
        self.preds = self.build_multires_prediction_layer()
        self.encoders = self.encoder_block(
            arc_type=self.arc_type, patch_embed_type=self.patch_embed_type, img_size=self.input_size,
            patch_size=self.patch_size, in_chans=self.num_bins_events, embed_dim=self.base_num_channels,
            depths=self.depths, num_heads=self.num_heads, window_size=self.window_size,
            pretrained_window_size=self.pretrained_window_size, mlp_ratio=self.mlp_ratio, out_indices=self.out_indices,
            norm=self.norm, spikformer_norm=self.spikformer_norm, pol_in_channel=self.pol_channel, **self.spiking_kwargs)
        # constructed-but-unused in the reference too (:154-156); kept for module-tree parity (no parameters)
        self.preds_out = nn.ModuleList([neuron.IFNode(v_threshold=float("inf"), v_reset=0.0)
                                        for _ in range(self.num_encoders)])

    def forward_cl(self, x):
        """voxels (B, bins, 2, H, W) -> multi-resolution predictions, each (B, T, Hi, Wi, 2) channels-last."""
        blocks = self.encoders.forward_cl(x)
        x = blocks[-1]
        for resblock in self.resblocks:
            x = resblock.forward_cl(x)
        predictions = []
        for i, (decoder, pred) in enumerate(zip(self.decoders, self.preds)):
            x = skip_concat_cl(x, blocks[self.num_encoders - i - 1])
            if i > 0:
                x = skip_concat_cl(predictions[-1], x, align=16)     # 16: the TMA pixel pitch of the 1-byte spike operand
            x = decoder.forward_cl(x)
            predictions.append(pred.forward_cl(x))
        return predictions

    def forward(self, x):
        """Reference layout: list of (T, B, 2, Hi, Wi)."""
        if self.skip_type != "concat":
            raise NotImplementedError("skip_type 'sum' is not built (FlowNets hard-code 'concat', STSwinNet.py:336)")
        return [p.permute(1, 0, 4, 2, 3) for p in self.forward_cl(x)]


class MS_Spikingformer_MultiResUNet(Spikingformer_MultiResUNet):
    """Same with membrane-potential (MS) shortcuts (reference :239-252)."""
    pol_channel = False
    encoder_block = MS_spiking_former_encoder
    ff_type = MS_SpikingConvEncoderLayer
    res_type = MS_ResBlock
    transpose_type = MS_SpikingTransposeDecoderLayer
    pred_type = MS_SpikingPredLayer
    w_scale_pred = 0.01


class SpikingformerFlowNet(nn.Module):
    """SEW shortcut, 3 encoders (reference :254-311)."""
    unet_type = Spikingformer_MultiResUNet
    recurrent_block_type = "none"
    spiking_feedforward_block_type = None
    num_en = 3

    def __init__(self, unet_kwargs, stt_kwargs):
        super().__init__()
        flownet_kwargs = {
            "base_num_channels": unet_kwargs["base_num_channels"], "num_encoders": self.num_en,
            "num_residual_blocks": 2, "num_output_channels": 2, "skip_type": "concat",
            "norm": unet_kwargs.get("norm", None), "use_upsample_conv": unet_kwargs.get("use_upsample_conv", True),
            "kernel_size": unet_kwargs["kernel_size"], "channel_multiplier": 2,
            "recurrent_block_type": self.recurrent_block_type, "final_activation": unet_kwargs["final_activation"],
            "spiking_feedforward_block_type": self.spiking_feedforward_block_type,
            "spiking_neuron": unet_kwargs["spiking_neuron"],
        }
        self.crop = None
        self.mask = unet_kwargs["mask_output"]
        self.norm_input = unet_kwargs.get("norm_input", False)
        self.encoding, self.num_bins = unet_kwargs["encoding"], unet_kwargs["num_bins"]
        self.num_encoders = flownet_kwargs["num_encoders"]
        self.num_split = self.num_bins // stt_kwargs["swin_patch_size"][0]
        self.final_activation = flownet_kwargs["final_activation"]
        unet_kwargs.update(flownet_kwargs)
        for k in ("name", "encoding", "round_encoding", "norm_input", "mask_output"):
            unet_kwargs.pop(k, None)
        self.sttmultires_unet = self.unet_type(unet_kwargs, stt_kwargs)

    def detach_states(self):
        pass

    def reset_states(self):
        pass

    def init_weights(self):
        """kaiming-normal Linear, xavier-uniform Conv2d, unit BatchNorm (reference :264-276)."""
        def _init(m):
            if isinstance(m, nn.Linear):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, (nn.LayerNorm, nn.BatchNorm2d)):
                nn.init.constant_(m.bias, 0)
                nn.init.constant_(m.weight, 1.0)
            elif isinstance(m, nn.Conv2d):
                nn.init.xavier_uniform_(m.weight)
        self.apply(_init)

    def forward(self, x, log=False):
        if log:
            raise NotImplementedError("log=True (attention-score dump) is not built; see Spiking_SwinTransformerBlock3D")
        H, W = x.shape[-2], x.shape[-1]
        flow_list = []
        # The reference trainer enters the model under torch.cuda.amp.autocast when a GradScaler is configured
        # (train_flow_parallel_supervised_SNN.py:248).  This path has nothing for autocast to down-cast: membranes integrate
        # in fp32 and the spike contractions are exact integer GEMMs, so autocast is switched off for the model's extent
        # (the scripts run unmodified; losses / GradScaler see fp32 flows, which autocast would have produced for the
        # final sum / interpolate anyway).
        with torch.autocast(device_type=x.device.type, enabled=False), ops.defer_nbt():
            x = x.float()
            for pred in self.sttmultires_unet.forward_cl(x):               # (B, T, h, w, 2)
                flow = torch.sum(pred, dim=1).permute(0, 3, 1, 2)           # sum over time -> (B, 2, h, w)
                flow_list.append(torch.nn.functional.interpolate(
                    flow, scale_factor=(H / flow.shape[-2], W / flow.shape[-1])))
        return {"flow": flow_list, "attn": None}

    def __str__(self):
        n = sum(p.numel() for p in self.parameters() if p.requires_grad)
        return super().__str__() + f"\nTrainable parameters: {n}"


class MS_SpikingformerFlowNet(SpikingformerFlowNet):
    """MS shortcut, 3 encoders."""
    unet_type = MS_Spikingformer_MultiResUNet


class MS_SpikingformerFlowNet_en4(SpikingformerFlowNet):
    """MS shortcut, 4 encoders (the shipped SDformerFlow model)."""
    unet_type = MS_Spikingformer_MultiResUNet
    num_en = 4
