"""``YOGO`` - same Python surface as /root/reference/yogo/model.py:13-313, executed by the
sm_100a kernel plan in ``yogo_b200.engine`` instead of ``self.model(x)`` + ATen/cuDNN.

Kept verbatim-compatible: constructor signature and defaults, the 11 registered buffers, the
``model.*`` parameter keys (state_dict / ``.pth`` drop-in, model.py:94-147), ``Sx/Sy``,
``from_pth``, ``get_grid_size``, ``resize_model``, ``num_params``/``grad_norm``/``param_norm``.
Differences, all deliberate:
  * ``forward`` requires CUDA tensors on a B200 and raises otherwise (no CPU fallback);
  * the per-parameter ``clamp(grad, +-clip_value)`` hooks (model.py:76-77) are fused into
    the weight-gradient kernels, so no tensor hooks are registered;
  * ``compute_dtype`` (torch.bfloat16 default, torch.float32 for the tight-parity path)
    selects activation storage; parameters stay fp32.
"""
from __future__ import annotations

import os
from pathlib import Path
from typing import Any, Dict, Optional, Tuple, Union

import torch
from torch import nn

from .model_defns import ModelDefn, base_model, get_model_func

PathLike = Union[Path, str]


def _default_dtype() -> torch.dtype:
    v = os.environ.get("YOGO_B200_DTYPE", "bf16").lower()
    return torch.float32 if v in ("fp32", "f32", "float32") else torch.bfloat16


class YOGO(nn.Module):
    def __init__(
        self,
        img_size: Tuple[int, int],
        anchor_w: float,
        anchor_h: float,
        num_classes: int,
        is_rgb: bool = False,
        normalize_images: bool = False,
        inference: bool = False,
        tuning: bool = False,
        model_func: ModelDefn = base_model,
        clip_value: float = 1.0,
        device: Union[torch.device, str] = "cpu",
    ):
        super().__init__()
        self.device = device
        self.model = model_func(num_classes, is_rgb).to(device)
        self.model_version = model_func.__name__

        self.register_buffer("img_size", torch.tensor(img_size))
        self.register_buffer("anchor_w", torch.tensor(anchor_w))
        self.register_buffer("anchor_h", torch.tensor(anchor_h))
        self.register_buffer("num_classes", torch.tensor(num_classes))
        self.register_buffer("clip_value", torch.tensor(clip_value))
        self.register_buffer("is_rgb", torch.tensor(is_rgb))
        self.register_buffer("normalize_images", torch.tensor(normalize_images))

        self.inference = inference
        self.compute_dtype: torch.dtype = _default_dtype()
        self._clip_value_f = float(clip_value)
        self._runner = None

        Sx, Sy = self.get_grid_size()
        self.Sx, self.Sy = Sx, Sy
        _Cxs, _Cys = self._grid_offsets(Sx, Sy, self.device)
        self.register_buffer("_Cxs", _Cxs)
        self.register_buffer("_Cys", _Cys)
        self.register_buffer("height_multiplier", torch.tensor(1.0))
        self.register_buffer("width_multiplier", torch.tensor(1.0))

        self._refresh_host_consts()
        if tuning:
            self.model.apply(self.set_bn_eval)  # model.py:69-70
        else:
            self.model.apply(self.init_network_weights)  # model.py:71-73

    def _refresh_host_consts(self) -> None:
        """Host-side copies of the scalar buffers the kernels take by value.  Reading them from the device
        buffers on every forward would be a device->host sync per step (and is illegal during CUDA-graph
        capture); they only change in __init__, load_state_dict and resize_model."""
        self._hc = {
            "anchor_w": float(self.anchor_w), "anchor_h": float(self.anchor_h),
            "width_multiplier": float(self.width_multiplier) if hasattr(self, "width_multiplier") else 1.0,
            "height_multiplier": float(self.height_multiplier) if hasattr(self, "height_multiplier") else 1.0,
        }
        # (the gradient clamp keeps the CONSTRUCTOR's clip_value, like the reference, whose hook closes over the argument
        # (model.py:76-77): a `clip_value` buffer loaded from a checkpoint does not change it)

    def load_state_dict(self, state_dict, *args, **kwargs):
        out = super().load_state_dict(state_dict, *args, **kwargs)
        self._refresh_host_consts()
        return out

    @staticmethod
    def _grid_offsets(Sx: int, Sy: int, device) -> Tuple[torch.Tensor, torch.Tensor]:
        # model.py:48-55
        cxs = torch.linspace(0, 1 - 1 / Sx, Sx).expand(Sy, -1).to(device)
        cys = torch.linspace(0, 1 - 1 / Sy, Sy).expand(1, -1).transpose(0, 1).expand(Sy, Sx).to(device)
        return cxs.clone(), cys.clone()

    @staticmethod
    def init_network_weights(module: nn.Module):
        # model.py:79-87: Kaiming-normal(fan_out, a=0.01) conv weights, zero biases
        if isinstance(module, nn.Conv2d):
            torch.nn.init.kaiming_normal_(module.weight, a=0.01, mode="fan_out", nonlinearity="leaky_relu")
            if module.bias is not None:
                torch.nn.init.zeros_(module.bias)

    @staticmethod
    def set_bn_eval(module: nn.Module):
        if isinstance(module, torch.nn.modules.batchnorm._BatchNorm):
            module.eval()

    @classmethod
    def from_pth(cls, pth_path: PathLike, inference: bool = False) -> Tuple["YOGO", Dict[str, Any]]:
        """model.py:94-147 - same checkpoint dict, same permissiveness for old files."""
        pth_path = Path(pth_path)
        loaded_pth = torch.load(pth_path, map_location="cpu")

        global_step = loaded_pth.get("step", 0)
        model_version = loaded_pth.get("model_version", None)
        class_names = loaded_pth.get("class_names", None)

        params = loaded_pth["model_state_dict"]
        img_size = params["img_size"]
        anchor_w = params["anchor_w"]
        anchor_h = params["anchor_h"]
        num_classes = params["num_classes"]

        for key, default in (
            ("is_rgb", torch.tensor(False)),
            ("clip_value", torch.tensor(1.0)),
            ("height_multiplier", torch.tensor(1.0)),
            ("width_multiplier", torch.tensor(1.0)),
        ):
            if key not in params:
                params[key] = default
        if "normalize_images" not in params:
            params["normalize_images"] = torch.tensor(loaded_pth.get("normalize_images", False))

        model = cls(
            (int(img_size[0]), int(img_size[1])),
            anchor_w.item(),
            anchor_h.item(),
            num_classes=num_classes.item(),
            inference=inference,
            tuning=True,
            model_func=get_model_func(model_version),
        )
        model.load_state_dict(params)
        if inference:
            model.eval()
        return model, {
            "step": global_step,
            "class_names": class_names,
            "normalize_images": params["normalize_images"],
        }

    def to(self, device, *args, **kwargs):
        self.device = device
        super().to(device, *args, **kwargs)
        return self

    def num_params(self) -> int:
        return sum(p.numel() for p in self.parameters() if p.requires_grad)

    def grad_norm(self) -> float:
        total = 0.0
        for p in self.parameters():
            if p.grad is not None and p.requires_grad:
                total += p.grad.detach().norm(2).item() ** 2
        return total**0.5

    def param_norm(self) -> float:
        total = 0.0
        for p in self.parameters():
            if p.grad is not None and p.requires_grad:
                total += p.detach().norm(2).item() ** 2
        return total**0.5

    def get_img_size(self) -> Tuple[torch.Tensor, torch.Tensor]:
        if isinstance(self.img_size, torch.Tensor):
            h, w = self.img_size
            return h, w
        raise ValueError(f"self.img_size is not a tensor: {type(self.img_size)}")

    def get_grid_size(self, img_size: Optional[Tuple[int, int]] = None) -> Tuple[int, int]:
        """return Sx, Sy (model.py:189-234): walk the conv modules with integer arithmetic."""
        if img_size is not None:
            h, w = int(img_size[0]), int(img_size[1])
        else:
            hh, ww = self.get_img_size()
            h, w = int(hh), int(ww)

        def pair(v) -> Tuple[int, int]:
            if isinstance(v, tuple):
                return int(v[0]), int(v[1])
            if v is None or v == "none":
                return 0, 0
            return int(v), int(v)

        for mod in self.modules():
            if isinstance(mod, nn.Conv2d):
                (p0, p1), (d0, d1), (k0, k1), (s0, s1) = (
                    pair(mod.padding), pair(mod.dilation), pair(mod.kernel_size), pair(mod.stride))
                h = (h + 2 * p0 - d0 * (k0 - 1) - 1) // s0 + 1
                w = (w + 2 * p1 - d1 * (k1 - 1) - 1) // s1 + 1
            elif isinstance(mod, nn.ConvTranspose2d):
                (p0, p1), (d0, d1), (k0, k1), (s0, s1) = (
                    pair(mod.padding), pair(mod.dilation), pair(mod.kernel_size), pair(mod.stride))
                o0, o1 = pair(mod.output_padding)
                h = (h - 1) * s0 - 2 * p0 + d0 * (k0 - 1) + o0 + 1
                w = (w - 1) * s1 - 2 * p1 + d1 * (k1 - 1) + o1 + 1
        return int(w), int(h)

    def resize_model(self, img_height: Optional[int] = None, img_width: Optional[int] = None) -> None:
        """model.py:236-265: re-grid for a cropped field of view."""
        org_h, org_w = (int(d) for d in self.get_img_size())
        crop = (img_height or org_h, img_width or org_w)
        Sx, Sy = self.get_grid_size(crop)
        self.Sx, self.Sy = Sx, Sy
        dev = self.img_size.device
        _Cxs, _Cys = self._grid_offsets(Sx, Sy, dev)
        self.register_buffer("height_multiplier", torch.tensor(org_h / crop[0], device=dev))
        self.register_buffer("width_multiplier", torch.tensor(org_w / crop[1], device=dev))
        self.register_buffer("img_size", torch.tensor(crop, device=dev))
        self.register_buffer("_Cxs", _Cxs)
        self.register_buffer("_Cys", _Cys)
        self._refresh_host_consts()

    def _get_runner(self):
        from . import engine

        if self._runner is None:
            self._runner = engine.Runner(self)
        return self._runner

    def set_fused_input_scale(self, scale: Optional[float]) -> None:
        """uint8 images are consumed as `x * scale` by the first-layer kernels (scale = 1/255 reproduces the `/ 255` that
        the reference's datasets apply when `normalize_images` is set, yogo_dataset.py:280-283) - no fp32 copy of the
        image is made.  `None` restores the reference's forward (`x.float()`, model.py:272-273).  Float inputs are never
        scaled."""
        self._get_runner().input_scale = scale

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """model.py:267-313.  Returns (N, 5+C, Sy, Sx) fp32: [xc, yc, w, h, objectness, classes...]."""
        from . import engine

        return engine.run(self._get_runner(), x)
