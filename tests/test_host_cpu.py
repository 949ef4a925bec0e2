"""CPU tests of the host side: C-ABI library loads and exports what include/sdf_b200.h declares,
ctypes layouts match the C compiler's, argument validation fails loudly, the product has no CPU
fallback, module surface / state_dict layout match the reference contract."""
import ctypes
import os
import shutil
import subprocess

import pytest
import torch

from oracle import synth


def test_library_exports_every_declared_symbol():
    from sdformerflow_b200 import capi
    L = capi.lib()
    names = capi.declared_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), n
    assert capi.version() == 1
    assert L.sdf_last_error() is not None


@pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc not available")
def test_ctypes_layouts_match_c(tmp_path):
    from sdformerflow_b200 import capi
    st, _ = capi.parse_header()
    src = f'#include "{capi.HEADER}"\n#include <stdio.h>\n#include <stddef.h>\nint main(){{\n'
    for k, v in st.items():
        src += f'printf("{k} %zu", sizeof({k}));\n'
        for f, _t in v["fields"]:
            src += f'printf(" %zu", offsetof({k},{f}));\n'
        src += 'printf("\\n");\n'
    src += "return 0;}\n"
    c = tmp_path / "sz.c"
    c.write_text(src)
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", str(c), "-o", str(exe)])
    for line in subprocess.check_output([str(exe)], text=True).splitlines():
        parts = line.split()
        cls = st[parts[0]]["ctype"]
        mine = [ctypes.sizeof(cls)] + [getattr(cls, f).offset for f, _ in st[parts[0]]["fields"]]
        assert mine == list(map(int, parts[1:])), parts[0]


def test_argument_validation_fails_loudly():
    from sdformerflow_b200 import capi
    with pytest.raises(RuntimeError, match="null argument"):
        capi.call("sdf_lif_fwd", capi.struct("sdf_lif_fwd_args"))
    with pytest.raises(RuntimeError, match="shift must be in"):
        capi.call("sdf_window_index", capi.struct("sdf_window_index_args", win2x=16,
                                                  g=dict(B=1, D=2, H=4, W=4, wd=2, wh=2, ww=2, sd=2, sh=0, sw=0)))
    g = capi.struct("sdf_window_geom", B=2, D=5, H=16, W=16, wd=2, wh=8, ww=8, sd=1, sh=4, sw=4)
    assert capi.lib().sdf_window_rows(ctypes.byref(g)) == 2 * 6 * 16 * 16


def test_no_cpu_fallback():
    from sdformerflow_b200 import ops, capi
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.neuron(torch.zeros(2, 8), ops.NeuronCfg())
    from helpers import build_product
    mc, sc = synth.small_config("lif")
    model = build_product(mc, sc, "cpu")
    with pytest.raises(RuntimeError):
        model(synth.synth_voxels(1, 10, 96, 128))


def test_product_never_imports_oracle():
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "sdformerflow_b200")
    for dp, _, files in os.walk(root):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dp, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, os.path.join(dp, f)


@pytest.mark.parametrize("nt,n_entries", [("lif", 501), ("psn", 711), ("plif", 606)])
def test_state_dict_layout_en4(nt, n_entries):
    """Appendix D of SURVEY.md: entry counts and key shapes of MS_SpikingformerFlowNet_en4."""
    from oracle import reference_loader as rl
    from helpers import product_template_sd
    mc, sc = rl.default_config(nt, input_size=(288, 384))
    sd = product_template_sd(mc, sc)
    assert len(sd) == n_entries
    p = "sttmultires_unet.encoders.swin3d."
    assert tuple(sd[p + "patch_embed.head.conv.0.weight"].shape) == (48, 2, 3, 3)
    assert tuple(sd[p + "layers.0.swin_blocks.0.attn.positional_encoding"].shape) == (1, 3, 162, 32)
    assert tuple(sd[p + "layers.2.swin_blocks.5.mlp.fc1.weight"].shape) == (1536, 384)
    assert tuple(sd[p + "layers.1.downsample.reduction.weight"].shape) == (384, 768)
    assert p + "layers.3.downsample.reduction.weight" not in sd
    assert tuple(sd["sttmultires_unet.preds.3.conv.0.bias"].shape) == (2,)
    if nt == "psn":
        assert tuple(sd[p + "layers.0.swin_blocks.0.attn.sn_q.spiking_neuron.weight"].shape) == (2, 2)
        assert tuple(sd[p + "layers.0.swin_blocks.0.mlp.sn1.spiking_neuron.bias"].shape) == (10, 1)
    n_params = sum(v.numel() for k, v in sd.items()
                   if not any(s in k for s in ("running_", "num_batches", "relative_position_index")))
    assert n_params == {"lif": 54913928, "psn": 54919238}.get(nt, n_params)


@pytest.mark.skipif(not os.path.isdir("/root/reference/models"), reason="reference tree only exists in the build container")
@pytest.mark.parametrize("nt", ["lif", "psn"])
def test_state_dict_keys_equal_reference(nt):
    from oracle import reference_loader as rl
    from helpers import product_template_sd
    mc, sc = rl.default_config(nt, input_size=(288, 384))
    ours = product_template_sd(mc, sc)
    ref = rl.build_reference_model(mc, sc).state_dict()
    assert list(ours.keys()) == list(ref.keys())
    for k in ref:
        assert ours[k].shape == ref[k].shape and ours[k].dtype == ref[k].dtype, k


def test_spikingjelly_protocol():
    from sdformerflow_b200.sj import functional, neuron
    from helpers import build_product
    mc, sc = synth.small_config("lif")
    model = build_product(mc, sc, "cpu")
    functional.reset_net(model)
    functional.set_step_mode(model, "m")
    functional.set_backend(model, "cupy", neuron.LIFNode)   # accepted no-op, as in the reference scripts
    lifs = [m for m in model.modules() if isinstance(m, neuron.LIFNode)]
    assert lifs and all(m.step_mode == "m" and m.backend == "cupy" for m in lifs)
    assert not any("spiking_neuron.v" in k for k in model.state_dict())


def test_window_geometry_host_math():
    from sdformerflow_b200 import ops
    g = ops.WindowGeom(8, 10, 120, 160, (2, 9, 9), (1, 4, 4))
    assert (g.Dp, g.Hp, g.Wp, g.nW, g.N, g.M) == (10, 126, 162, 1260, 162, 8 * 1260)
    g = ops.WindowGeom(1, 10, 9, 12, (2, 9, 9), (1, 4, 4))      # 288x384 stage 4: shift_h clamps to 0
    assert g.window == (2, 9, 9) and g.shift == (1, 0, 4) and g.nW == 5 * 1 * 2
    g = ops.WindowGeom(1, 5, 8, 8, (2, 8, 8), (1, 4, 4))        # MDR stage 4
    assert g.shift == (1, 0, 0) and g.Dp == 6


@pytest.mark.parametrize("bins,steps", [(10, 10), (20, 10), (10, 5), (12, 4)])
def test_bins_to_steps_regroup_matches_reference_loop(bins, steps):
    """The one-permute regroup == the reference's zero-fill + per-channel copy loop (Spiking_modules.py:1775-1786)."""
    from oracle import port
    from sdformerflow_b200.STSwinNet_SNN import Spiking_modules as m
    x = torch.randn(2, bins, 2, 6, 8)
    ref = port.regroup_events(x, bins, steps)
    assert torch.equal(ref, m.regroup_bins_to_steps(x, bins, steps))
    assert torch.equal(m.to_cl(ref), m.regroup_bins_to_steps_cl(x, bins, steps))


def test_deferred_batchnorm_counters():
    """ops.defer_nbt: the num_batches_tracked += 1 of every BatchNorm call of a forward pass becomes one multi-tensor add on
    exit — same counts as nn.BatchNorm2d's per-call increment, also for a module called more than once, nested scopes flush
    once (at the outermost exit)."""
    from sdformerflow_b200 import ops
    a, b, c = (torch.zeros((), dtype=torch.long) for _ in range(3))
    assert ops.NBT_DEFER is None
    with ops.defer_nbt():
        ops.NBT_DEFER.extend([a, b, b])
        with ops.defer_nbt():
            ops.NBT_DEFER.append(c)
        assert int(c) == 0                      # the inner scope does not flush
    assert (int(a), int(b), int(c)) == (1, 2, 1) and ops.NBT_DEFER is None
    with pytest.raises(ZeroDivisionError):      # an exception inside still leaves the switch off
        with ops.defer_nbt():
            1 / 0
    assert ops.NBT_DEFER is None


def test_spikes_take_grad_drops_the_token():
    """Spikes.take_grad (called by the producer's backward) returns the summed consumer gradients and releases the token:
    the producer node -> holder -> token -> grad_fn cycle must not keep a step's spikes alive until a cyclic GC pass."""
    from sdformerflow_b200 import ops
    s = ops.Spikes()
    s.data = torch.zeros(4, 8, dtype=torch.uint8)
    s.token = torch.zeros(())
    s.add_grad(torch.ones(4, 8))
    s.add_grad(torch.full((4, 8), 2.0))
    g = s.take_grad()
    assert torch.equal(g, torch.full((4, 8), 3.0)) and s.token is None and s.grad is None
    z = s.take_grad()                           # no consumer produced a gradient: zeros of the spike shape
    assert z.shape == (4, 8) and not z.any()


def test_pack_plan_recording_is_per_parameter():
    """gemm.start_recording / stop_recording (what train.GraphedStep builds its PackPlan from): parameters only, one job per
    (parameter, layout / padded width), in first-use order; nothing is recorded outside a recording."""
    from sdformerflow_b200 import gemm
    w1, w2 = torch.nn.Parameter(torch.zeros(8, 16)), torch.nn.Parameter(torch.zeros(16, 8, 3, 3))
    assert gemm._record is None and gemm.stop_recording() == []
    gemm.start_recording()
    gemm._record.append(("w", w1, "linear", True))
    gemm._record.append(("deconv", w2, 16))
    gemm._record.append(("w", w1, "linear", True))
    jobs = gemm.stop_recording()
    assert [j[0] for j in jobs] == ["w", "deconv"] and jobs[0][1] is w1 and jobs[1][1] is w2
    assert gemm._record is None and not gemm._pinned


def test_deconv_parity_class_taps_reproduce_conv_transpose():
    """sdf_spike_deconv_class_taps (host logic shared by the transposed-conv forward, its weight gradient and the stride-2 data
    gradient): output pixels of parity (a, b) = a stride-1 convolution of the input with the listed taps at shifts (dh, dw) in
    {0, 1}, rows / columns past the image reading zeros.  Rebuilt here with plain tensor ops and compared with
    F.conv_transpose2d(k 3, stride 2, padding 1, output_padding 1) — and, the same identity read backwards, with the data
    gradient of a 3x3 / stride-2 / padding-1 convolution."""
    import torch.nn.functional as F
    from sdformerflow_b200 import capi
    L = capi.lib()
    g = torch.Generator().manual_seed(11)
    N, H, W, Ci, Co = 2, 5, 7, 3, 4
    x = torch.randn(N, Ci, H, W, generator=g, dtype=torch.float64)
    w = torch.randn(Ci, Co, 3, 3, generator=g, dtype=torch.float64)
    out = torch.zeros(N, Co, 2 * H, 2 * W, dtype=torch.float64)
    xp = F.pad(x, (0, 1, 0, 1))                     # the zero fill past the last row / column
    seen = []
    for cls in range(4):
        src, dh, dw = ((ctypes.c_int64 * 4)() for _ in range(3))
        n = int(L.sdf_spike_deconv_class_taps(cls, src, dh, dw))
        a, b = cls >> 1, cls & 1
        assert n == (2 if a else 1) * (2 if b else 1)
        for t in range(n):
            kh, kw = int(src[t]) // 3, int(src[t]) % 3
            seen.append(int(src[t]))
            assert dh[t] in (0, 1) and dw[t] in (0, 1)
            xs = xp[:, :, dh[t]:dh[t] + H, dw[t]:dw[t] + W]
            out[:, :, a::2, b::2] += torch.einsum("nihw,io->nohw", xs, w[:, :, kh, kw])
    assert sorted(seen) == list(range(9))           # every kernel tap belongs to exactly one class
    ref = F.conv_transpose2d(x, w, None, stride=2, padding=1, output_padding=1)
    assert torch.allclose(out, ref, atol=1e-12)
    # the data gradient of conv2d(3x3, stride 2, padding 1) with weight w2 (Co2, Ci2, 3, 3) is that transposed convolution of
    # the output gradient with w2 read as (in = Co2, out = Ci2)
    w2 = torch.randn(Co, Ci, 3, 3, generator=g, dtype=torch.float64)
    xin = torch.zeros(N, Ci, 2 * H, 2 * W, dtype=torch.float64, requires_grad=True)
    gy = torch.randn(N, Co, H, W, generator=g, dtype=torch.float64)
    (gx,) = torch.autograd.grad(F.conv2d(xin, w2, None, stride=2, padding=1), xin, gy)
    assert torch.allclose(gx, F.conv_transpose2d(gy, w2, None, stride=2, padding=1, output_padding=1), atol=1e-12)
