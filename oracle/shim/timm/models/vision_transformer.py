"""timm.models.vision_transformer (0.5.4): only `_init_vit_weights`, which the reference's Swin TaskPrompter applies
to every sub-module at construction (TaskPrompter/models/transformers/taskprompter_swin.py:657-664). Published timm
0.5.4 behaviour for the non-jax path: Linear -> trunc_normal(std .02) weight, zero bias (the `head` / `pre_logits`
special cases do not occur in the reference); LayerNorm / GroupNorm / BatchNorm2d -> zero bias, unit weight; Conv2d is
left at PyTorch's default in the non-jax path."""
import torch.nn as nn

from .layers import lecun_normal_, trunc_normal_


def _init_vit_weights(module: nn.Module, name: str = '', head_bias: float = 0., jax_impl: bool = False):
    if isinstance(module, nn.Linear):
        if jax_impl:
            nn.init.xavier_uniform_(module.weight)
        else:
            trunc_normal_(module.weight, std=.02)
        if module.bias is not None:
            nn.init.zeros_(module.bias)
    elif jax_impl and isinstance(module, nn.Conv2d):
        lecun_normal_(module.weight)
        if module.bias is not None:
            nn.init.zeros_(module.bias)
    elif isinstance(module, (nn.LayerNorm, nn.GroupNorm, nn.BatchNorm2d)):
        nn.init.zeros_(module.bias)
        nn.init.ones_(module.weight)
