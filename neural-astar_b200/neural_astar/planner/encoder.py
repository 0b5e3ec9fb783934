"""Cost-map encoders (plain PyTorch; the north star keeps them on cuDNN).

Same classes, constructor arguments and state-dict layout as
/root/reference/src/neural_astar/planner/encoder.py:12-97 so reference checkpoints load:
`model.{0,3,6,...}` are 3x3 convs, `model.{1,4,7,...}` BatchNorms, ReLU (and MaxPool for
CNNDownSize) in between, trailing activation(s) dropped; output = sigmoid(model(x)) * const.
"""
from __future__ import annotations

import os
from typing import List, Optional

import torch
import torch.nn as nn

# Precision of the eval fast path's 3x3 convolutions: TF32 tensor-core math (what torch's own default
# `torch.backends.cudnn.allow_tf32 = True` gives the reference on any Ampere+ GPU) or full fp32.  The head
# (last, single-output-channel conv) is always fp32.  bench.py states this in its `dtype` field;
# tests/test_gpu_module_api.py counts how many search masks differ between the two settings.
ALLOW_TF32 = os.environ.get("NASTAR_B200_ENCODER_TF32", "1") != "0"
# The head's per-pixel [C] x [C,9] product: the engine's streaming kernel (csrc/nastar_glue.cuh head_taps_kernel, fp32
# FMA, weights in the constant bank) or, with NASTAR_B200_HEAD_KERNEL=0, torch.mm (cuBLAS SGEMM, 3-4x slower on this
# skinny shape).  Same sums up to fp32 re-association.
HEAD_KERNEL = os.environ.get("NASTAR_B200_HEAD_KERNEL", "1") != "0"
# First layer of the "m+" CNN (2 -> 32 channels, 18 multiply-adds per output): the engine's kernel that also forms the
# start+goal channel and the concat on the fly (csrc/nastar_glue.cuh conv1_marks_kernel, fp32 FMA), or, with
# NASTAR_B200_CONV1_KERNEL=0, the pack_inputs kernel followed by cuDNN's (also fp32, generic NHWC) convolution.
CONV1_KERNEL = os.environ.get("NASTAR_B200_CONV1_KERNEL", "1") != "0"


class EncoderBase(nn.Module):
    def __init__(self, input_dim: int, encoder_depth: int = 4, const: Optional[float] = None):
        super().__init__()
        self.model = self.construct_encoder(input_dim, encoder_depth)
        # learnable scale when given (reference :25-28); plain 1.0 otherwise
        self.const = nn.Parameter(torch.ones(1) * const) if const is not None else 1.0
        self._plan = None       # cached inference plan (folded weights) for the cuDNN fast path
        self._plan_key = None
        self._nhwc = False      # conv weights converted to channels-last (done lazily on the first CUDA batch)
        self._head_scalars = (0.0, 1.0)   # (folded bias of the head conv, const) as Python floats
        self._head_w_host = None          # folded [C,9] weights of the head conv on the host (kernel parameter)
        self._conv1 = None                # folded ([9,2,32] weights, [32] bias) of a 2->32 first layer, NumPy on the host

    def construct_encoder(self, input_dim: int, encoder_depth: int) -> nn.Module:
        raise NotImplementedError

    def fast_path_ok(self, x: torch.Tensor) -> bool:
        """Eval-mode, no-grad, fp32 CUDA input: the folded cuDNN plan (and the fused hand-off) may be used."""
        return (not self.training) and x.is_cuda and not torch.is_grad_enabled() and x.dtype == torch.float32

    def prepare_channels_last(self) -> None:
        """One-off, explicit: store the conv weights channels-last (cuDNN's tensor-core convolutions are
        NHWC-native; training step 3.45 -> 2.42 ms on B200).  Pure memory-format change: the Parameter objects,
        shapes and state-dict contents are unchanged.  Call it once after moving the module to the GPU — before
        wrapping it in DDP or capturing a CUDA graph.  forward() calls it on the first CUDA batch otherwise."""
        if not self._nhwc and isinstance(self.model, nn.Sequential):
            self.model.to(memory_format=torch.channels_last)
            self._nhwc = True

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.fast_path_ok(x):
            head = self.head_taps(x)
            if head is not None:
                # sigmoid(.)*const finished by the engine's glue kernel: the very arithmetic the fused search
                # prologue performs (NASTAR_COST_TAPS), so encode() and the fused forward agree bit for bit
                from .. import _native

                taps, bias, scale = head
                return _native.cost_from_taps(taps, bias, scale)
            plan = self._inference_plan(x.device)
            if plan is not None:
                return torch.sigmoid(_run_plan(plan, x)) * self.const
        if x.is_cuda and x.dim() == 4 and isinstance(self.model, nn.Sequential):
            self.prepare_channels_last()
            x = x.contiguous(memory_format=torch.channels_last)
        return torch.sigmoid(self.model(x)) * self.const

    def head_taps(self, x: torch.Tensor, out: Optional[torch.Tensor] = None, on_last_conv=None):
        """Eval fast path up to (not including) the 9-tap gather of the single-output-channel head:
        returns (taps [B,H,W,9] fp32 contiguous, folded bias, const) as (tensor, float, float), or None when the
        encoder does not end in such a head.  bias/const are cached Python floats (no host sync per call)."""
        plan = self._inference_plan(x.device)
        if plan is None or plan[-1][7] is None:
            return None
        x = x.contiguous(memory_format=torch.channels_last)
        return self._head_from(plan, 0, x, out, on_last_conv)

    def head_taps_marks(self, map_designs: torch.Tensor, start_maps: torch.Tensor, goal_maps: torch.Tensor,
                        out: Optional[torch.Tensor] = None, on_last_conv=None):
        """head_taps() of the "m+" input cat(map_designs, start_maps + goal_maps) without materialising it: the first
        layer reads the three planes directly (one kernel instead of pack_inputs + cuDNN's conv1).  None when the
        encoder's first layer is not the 2->32 3x3/ReLU block, the maps are not one-channel or the marks live on a
        different grid; callers then pack the input and use head_taps()."""
        if not CONV1_KERNEL or map_designs.shape[1] != 1 or map_designs.shape[-2:] != start_maps.shape[-2:]:
            return None
        plan = self._inference_plan(map_designs.device)
        if plan is None or plan[-1][7] is None or self._conv1 is None or len(plan) < 2:
            return None
        from .. import _native

        x = _native.conv1_marks(map_designs, start_maps, goal_maps, self._conv1[0], self._conv1[1])
        return self._head_from(plan, 1, x, out, on_last_conv)

    def _head_from(self, plan, first: int, x: torch.Tensor, out: Optional[torch.Tensor], on_last_conv=None):
        """Layers plan[first:-1] (cuDNN) and the head's products.  `on_last_conv`, if given, is called once, after the
        second-to-last convolution has been enqueued and before the last (widest) one: PipelinedPlanner forks the
        previous batch's search there (see its docstring)."""
        body = plan[first:-1]
        with _conv_flags():
            x = _run_plan_inner(body[:-1], x)
            if on_last_conv is not None:
                on_last_conv()
            x = _run_plan_inner(body[-1:], x)
        if not x.is_contiguous(memory_format=torch.channels_last):
            x = x.contiguous(memory_format=torch.channels_last)
        wm = plan[-1][7][0]
        B, C, H, W = x.shape
        if HEAD_KERNEL and self._head_w_host is not None and C in (32, 64, 128, 256):
            from .. import _native

            return _native.head_taps(x, self._head_w_host, out=out), self._head_scalars[0], self._head_scalars[1]
        a = x.permute(0, 2, 3, 1).reshape(-1, C)
        if out is not None:     # write the GEMM result straight into a caller-owned [B,H,W,9] buffer
            taps = torch.mm(a, wm, out=out.view(-1, 9)).view(B, H, W, 9)
        else:
            taps = torch.mm(a, wm).view(B, H, W, 9)
        return taps, self._head_scalars[0], self._head_scalars[1]

    # ---- eval-mode cuDNN fast path (encoder -> search hand-off) -----------------
    # Still plain PyTorch/cuDNN: BatchNorm (running stats) is folded into the preceding conv, activations
    # run channels-last (no NCHW<->NHWC transposes), and conv+bias+ReLU is one cuDNN fused op.  Same math
    # as self.model(x) up to fp32 re-association of the BN scale; training / autograd use self.model.
    def _inference_plan(self, device):
        if not isinstance(self.model, nn.Sequential):
            return None
        tensors = [t for t in list(self.model.parameters()) + list(self.model.buffers())]
        if isinstance(self.const, torch.Tensor):
            tensors.append(self.const)
        key = (str(device),) + tuple((t.data_ptr(), t._version) for t in tensors)
        if key == self._plan_key:
            return self._plan
        mods = list(self.model)
        plan, i = [], 0
        while i < len(mods):
            conv = mods[i]
            if not (isinstance(conv, nn.Conv2d) and i + 1 < len(mods) and isinstance(mods[i + 1], nn.BatchNorm2d)
                    and conv.groups == 1 and conv.bias is not None):
                self._plan, self._plan_key = None, key
                return None
            bn = mods[i + 1]
            i += 2
            relu = i < len(mods) and isinstance(mods[i], nn.ReLU)
            i += 1 if relu else 0
            pool = mods[i] if i < len(mods) and isinstance(mods[i], nn.MaxPool2d) else None
            i += 1 if pool is not None else 0
            with torch.no_grad():
                scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
                w = (conv.weight * scale.view(-1, 1, 1, 1)).contiguous(memory_format=torch.channels_last)
                b = ((conv.bias - bn.running_mean) * scale + bn.bias).contiguous()
            head = None
            if (conv.out_channels == 1 and conv.kernel_size == (3, 3) and conv.stride == (1, 1)
                    and conv.padding == (1, 1) and conv.dilation == (1, 1) and not relu and pool is None):
                # single-output-channel head: cuDNN's implicit GEMM wastes a 256-wide tile on one channel
                # (134 us at 100x256x32x32); a per-pixel [C] x [C,9] GEMM followed by a 9-tap gather is the
                # same sum in a different order (86 us) — see _conv3x3_single_output
                with torch.no_grad():
                    taps = torch.zeros(1, 9, 3, 3, device=w.device, dtype=w.dtype)
                    for k in range(9):
                        taps[0, k, k // 3, k % 3] = 1.0
                    head = (w[0].reshape(w.shape[1], 9).contiguous(), taps)
            plan.append((w, b, conv.stride, conv.padding, conv.dilation, relu, pool, head))
        # scalars of the fused hand-off, read back once per weight version (one host sync at plan build)
        last_bias = plan[-1][1]
        self._head_scalars = (float(last_bias.reshape(-1)[0]) if last_bias.numel() == 1 else 0.0,
                              float(self.const) if not isinstance(self.const, float) else self.const)
        head = plan[-1][7]
        self._head_w_host = head[0].detach().cpu().numpy().astype("float32", order="C") if head is not None else None
        w0, b0, st0, pad0, dil0, relu0, pool0, head0 = plan[0]
        self._conv1 = None
        if (tuple(w0.shape) == (32, 2, 3, 3) and tuple(st0) == (1, 1) and tuple(pad0) == (1, 1) and tuple(dil0) == (1, 1)
                and relu0 and pool0 is None and head0 is None and w0.is_cuda and w0.dtype == torch.float32):
            with torch.no_grad():       # [cout, cin, ky, kx] -> [tap, cin, cout]
                self._conv1 = (w0.permute(2, 3, 1, 0).reshape(9, 2, 32).cpu().numpy().astype("float32", order="C"),
                               b0.detach().cpu().numpy().astype("float32", order="C"))
        self._plan, self._plan_key = plan, key
        return plan


def _conv_flags():
    """cuDNN settings of the eval fast path.  Inference shapes are static, so cuDNN may time its candidates once
    per shape instead of trusting the heuristic (which picks a 2x slower kernel for the 64->128 layer on sm_100) —
    unless the user asked for determinism (set_global_seeds(), utils/training.py): autotuned algorithm choice can
    differ run to run, so `torch.backends.cudnn.deterministic = True` is honoured and turns autotuning off."""
    det = bool(torch.backends.cudnn.deterministic)
    return torch.backends.cudnn.flags(enabled=True, benchmark=not det, deterministic=det, allow_tf32=ALLOW_TF32)


def _run_plan(plan, x: torch.Tensor) -> torch.Tensor:
    x = x.contiguous(memory_format=torch.channels_last)
    with _conv_flags():
        return _run_plan_inner(plan, x)


def _conv3x3_single_output(x: torch.Tensor, wm: torch.Tensor, taps: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """3x3 / pad 1 convolution with ONE output channel on a channels-last activation:
    t[n,y,x,k] = <x[n,y,x,:], w[:,k]> for the 9 taps (one skinny fp32 GEMM over the pixels), then
    out[n,y,x] = b + sum_k t[n, y+ky-1, x+kx-1, k] (a 9->1 one-hot 3x3 convolution, kept in full fp32)."""
    B, C, H, W = x.shape
    t = (x.permute(0, 2, 3, 1).reshape(-1, C) @ wm).view(B, H, W, 9).permute(0, 3, 1, 2)
    det = bool(torch.backends.cudnn.deterministic)
    with torch.backends.cudnn.flags(enabled=True, benchmark=not det, deterministic=det, allow_tf32=False):
        return torch.nn.functional.conv2d(t, taps, b, padding=1)


def _run_plan_inner(plan, x: torch.Tensor) -> torch.Tensor:
    for w, b, stride, padding, dilation, relu, pool, head in plan:
        if head is not None and x.is_contiguous(memory_format=torch.channels_last):
            x = _conv3x3_single_output(x, head[0], head[1], b)
            continue
        if relu:
            x = torch.cudnn_convolution_relu(x, w, b, stride, padding, dilation, 1)
        else:
            x = torch.nn.functional.conv2d(x, w, b, stride, padding, dilation, 1)
        if pool is not None:
            x = pool(x)
    return x


def _conv_stack(widths: List[int], pool: bool) -> nn.Sequential:
    layers: List[nn.Module] = []
    tail = 2 if pool else 1  # modules to drop after the last BatchNorm
    for cin, cout in zip(widths[:-1], widths[1:]):
        layers += [nn.Conv2d(cin, cout, kernel_size=3, stride=1, padding=1), nn.BatchNorm2d(cout), nn.ReLU()]
        if pool:
            layers.append(nn.MaxPool2d((2, 2)))
    return nn.Sequential(*layers[:-tail])


class CNN(EncoderBase):
    CHANNELS = [32, 64, 128, 256]

    def construct_encoder(self, input_dim: int, encoder_depth: int) -> nn.Module:
        return _conv_stack([input_dim] + self.CHANNELS[:encoder_depth] + [1], pool=False)


class CNNDownSize(CNN):
    def construct_encoder(self, input_dim: int, encoder_depth: int) -> nn.Module:
        return _conv_stack([input_dim] + self.CHANNELS[:encoder_depth] + [1], pool=True)


class Unet(EncoderBase):
    """U-Net encoder; needs the optional third-party `segmentation_models_pytorch` (reference :37-57)."""

    DECODER_CHANNELS = [256, 128, 64, 32, 16]

    def construct_encoder(self, input_dim: int, encoder_depth: int) -> nn.Module:
        try:
            import segmentation_models_pytorch as smp
        except ImportError as e:  # not shipped in this image; used by no reference config
            raise ImportError("encoder_arch='Unet' needs segmentation_models_pytorch==0.3.1") from e
        return smp.Unet(
            encoder_name="vgg16_bn",
            encoder_weights=None,
            classes=1,
            in_channels=input_dim,
            encoder_depth=encoder_depth,
            decoder_channels=self.DECODER_CHANNELS[:encoder_depth],
        )
