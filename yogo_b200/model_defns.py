"""Backbone registry: the plug-in point of the reference (/root/reference/yogo/model_defns.py:6-27).

A ``ModelDefn`` is ``Callable[[int, bool], nn.Module]`` returning an ``nn.Sequential`` of conv
blocks.  The modules built here are ordinary ``torch.nn`` modules, so ``state_dict`` keys,
shapes and dtypes are identical to the reference's and user-registered
definitions keep working; they are never *called* - ``yogo_b200.engine`` compiles the
Sequential into a plan of sm_100a kernels (and raises for modules it cannot express).

Every definition of the reference is generated from a small table instead of being written
out block by block; the line ranges cite the reference definition each row restates.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Tuple

from torch import nn

ModelDefn = Callable[[int, bool], nn.Module]

MODELS: Dict[str, ModelDefn] = {}


def get_model_func(model_name: Optional[str]) -> ModelDefn:
    """Unknown or missing names silently map to ``base_model`` (model_defns.py:11-18)."""
    if model_name is None:
        return base_model
    return MODELS.get(model_name, base_model)


def register_model(model_defn: ModelDefn) -> ModelDefn:
    MODELS[model_defn.__name__] = model_defn
    return model_defn


# row = (out_channels, stride, conv_bias, batchnorm, dropout_p)
_Row = Tuple[int, int, bool, bool, float]


def _make_act(kind: str) -> nn.Module:
    return nn.SiLU(inplace=True) if kind == "silu" else nn.LeakyReLU()


def _build(rows: List[_Row], num_classes: int, rgb_input: bool, act: str = "lrelu") -> nn.Sequential:
    cin = 3 if rgb_input else 1
    blocks: List[nn.Module] = []
    for cout, stride, bias, bn, p in rows:
        mods: List[nn.Module] = [nn.Conv2d(cin, cout, 3, stride=stride, padding=1, bias=bias)]
        if bn:
            mods.append(nn.BatchNorm2d(cout))
        mods.append(_make_act(act))
        if p > 0:
            mods.append(nn.Dropout2d(p=p))
        blocks.append(nn.Sequential(*mods))
        cin = cout
    blocks.append(nn.Conv2d(cin, 5 + num_classes, 1))
    return nn.Sequential(*blocks)


def _eight(c1: int, c2: int, c3: int, c4: int) -> List[_Row]:
    return [
        (c1, 2, False, True, 0.0),
        (c2, 1, True, False, 0.05),
        (c3, 2, True, False, 0.10),
        (c4, 1, True, False, 0.15),
        (c4, 2, False, True, 0.0),
        (c4, 1, True, True, 0.0),
        (c4, 1, True, False, 0.0),
    ]


@register_model
def base_model(num_classes: int, rgb_input: bool = False) -> nn.Module:  # model_defns.py:30-77
    return _build(_eight(16, 32, 64, 128), num_classes, rgb_input)


@register_model
def silu_model(num_classes: int, rgb_input: bool = False) -> nn.Module:  # :80-127
    return _build(_eight(16, 32, 64, 128), num_classes, rgb_input, act="silu")


@register_model
def double_filters(num_classes: int, rgb_input: bool = False) -> nn.Module:  # :130-177
    return _build(_eight(32, 64, 128, 256), num_classes, rgb_input)


@register_model
def triple_filters(num_classes: int, rgb_input: bool = False) -> nn.Module:  # :180-227
    return _build(_eight(48, 96, 192, 384), num_classes, rgb_input)


@register_model
def half_filters(num_classes: int, rgb_input: bool = False) -> nn.Module:  # :230-277
    return _build(_eight(8, 16, 32, 64), num_classes, rgb_input)


@register_model
def quarter_filters(num_classes: int, rgb_input: bool = False) -> nn.Module:  # :280-327
    return _build(_eight(4, 8, 16, 32), num_classes, rgb_input)


@register_model
def depth_ver_0(num_classes: int, rgb_input: bool = False) -> nn.Module:  # :330-355
    rows = [(32, 2, False, True, 0.0), (128, 2, True, False, 0.10), (128, 2, False, True, 0.0)]
    return _build(rows, num_classes, rgb_input)


@register_model
def depth_ver_1(num_classes: int, rgb_input: bool = False) -> nn.Module:  # :358-392
    rows = [
        (16, 2, False, True, 0.0),
        (64, 2, True, False, 0.10),
        (128, 1, True, False, 0.15),
        (128, 2, False, True, 0.0),
        (128, 1, True, False, 0.0),
    ]
    return _build(rows, num_classes, rgb_input)


@register_model
def depth_ver_2(num_classes: int, rgb_input: bool = False) -> nn.Module:  # :395-397
    return base_model(num_classes, rgb_input)


@register_model
def depth_ver_3(num_classes: int, rgb_input: bool = False) -> nn.Module:  # :400-459
    rows = [
        (16, 2, False, True, 0.0),
        (32, 1, True, False, 0.05),
        (32, 1, True, False, 0.05),
        (64, 2, True, False, 0.10),
        (128, 1, True, False, 0.15),
        (128, 1, True, True, 0.0),
        (128, 2, False, False, 0.0),  # stride 2 without BatchNorm (:433-436)
        (128, 1, True, True, 0.0),
        (128, 1, True, False, 0.0),
    ]
    return _build(rows, num_classes, rgb_input)


@register_model
def depth_ver_4(num_classes: int, rgb_input: bool = False) -> nn.Module:  # :462-529
    rows = [
        (16, 2, False, True, 0.0),
        (16, 1, True, False, 0.0),
        (32, 1, True, False, 0.05),
        (32, 1, True, False, 0.05),
        (64, 2, True, False, 0.10),
        (64, 1, True, False, 0.0),
        (128, 1, True, False, 0.15),
        (128, 1, True, True, 0.0),
        (128, 2, True, False, 0.0),  # stride 2 with bias, no BatchNorm (:502-505)
        (128, 1, True, True, 0.0),
        (128, 1, True, False, 0.0),
    ]
    return _build(rows, num_classes, rgb_input)


# `convnext_small` (model_defns.py:532-558) needs timm and ConvTranspose2d; it is outside the
# hot-path scope and is intentionally not registered here.
