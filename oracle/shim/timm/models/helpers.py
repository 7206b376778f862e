"""timm.models.helpers stand-ins. build_model_with_cfg ignores `pretrained` (the reference hard-codes
pretrained=True, TaskPrompter/utils/common_config.py:22, which would need the network)."""
import torch.nn as nn


def build_model_with_cfg(model_cls, variant, pretrained, default_cfg=None, **kwargs):
    for k in ("pretrained_filter_fn", "pretrained_custom_load", "feature_cfg", "pretrained_strict"):
        kwargs.pop(k, None)
    model = model_cls(**kwargs)
    model.default_cfg = default_cfg
    return model


def named_apply(fn, module: nn.Module, name='', depth_first=True, include_root=False):
    if not depth_first and include_root:
        fn(module=module, name=name)
    for child_name, child in module.named_children():
        child_name = '.'.join((name, child_name)) if name else child_name
        named_apply(fn=fn, module=child, name=child_name, depth_first=depth_first, include_root=True)
    if depth_first and include_root:
        fn(module=module, name=name)
    return module


def adapt_input_conv(in_chans, conv_weight):
    return conv_weight


def overlay_external_default_cfg(default_cfg, kwargs):
    return default_cfg
