"""Imports the UNMODIFIED reference models from /root/reference through the timm/easydict shim.

TEST INFRASTRUCTURE. /root/reference exists only in the build container (not on the GPU box), so
everything here is used to (a) pin oracle/*_ref.py against the real reference and (b) generate the
golden vectors in tests/golden/. The two reference roots (TaskPrompter/, InvPT/) both define
top-level packages `models`, `utils`, ... so only one can be active at a time: switching purges
the other's modules from sys.modules.
"""
import importlib
import os
import sys

_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shim")
_COLLIDING = ("models", "utils", "data", "evaluation", "losses", "configs", "detection_toolbox")


_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def reference_root():
    """$MTT_REFERENCE, /root/reference (the build container), then <repo>/baseline/_ref (a copy of the reference tree
    that travels to the GPU box when someone puts one there; the reference has no installable package)."""
    for cand in (os.environ.get("MTT_REFERENCE"), "/root/reference", os.path.join(_REPO, "baseline", "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "TaskPrompter")):
            return cand
    return None


def available():
    return reference_root() is not None


def _activate(sub):
    root = os.path.join(reference_root(), sub)
    for k in list(sys.modules):
        if k.split(".")[0] in _COLLIDING:
            del sys.modules[k]
    sys.path[:] = [p for p in sys.path if not p.rstrip("/").endswith(("/TaskPrompter", "/InvPT"))]
    if _SHIM not in sys.path:
        sys.path.insert(0, _SHIM)
    sys.path.insert(0, root)
    importlib.invalidate_caches()


def build_taskprompter(cfg):
    """Reference TaskPrompterWrapper for a config dict from oracle.configs.taskprompter()."""
    import torch.nn as nn

    _activate("TaskPrompter")
    from easydict import EasyDict
    from models.transformers.taskprompter import TaskPrompter, ConvHead, DEConvHead
    from models.taskprompter_wrapper import TaskPrompterWrapper

    p = EasyDict(TASKS=EasyDict(NAMES=list(cfg["tasks"]), NUM_OUTPUT=dict(cfg["num_output"])),
                 prompt_len=cfg["prompt_len"], chan_nheads=cfg["chan_nheads"], use_ctr=cfg["use_ctr"],
                 embed_dim=cfg["e"], final_embed_dim=cfg["f"])
    backbone = TaskPrompter(p=p, select_list=list(cfg["select"]), img_size=tuple(cfg["img_size"]),
                            patch_size=cfg["patch"], embed_dim=cfg["C"], depth=cfg["depth"],
                            num_heads=cfg["heads"], chan_nheads=cfg["chan_nheads"], drop_path_rate=0.15)
    head_cls = DEConvHead if cfg.get("head", "conv") == "deconv" else ConvHead       # utils/common_config.py:64-70
    heads = nn.ModuleDict({t: head_cls(cfg["f"], cfg["num_output"][t]) for t in cfg["tasks"]})
    return TaskPrompterWrapper(p, backbone, heads)


def build_taskprompter_swin(cfg):
    """Reference TaskPrompterWrapper around TaskPrompterSwin for a dict from oracle.configs.taskprompter_swin()
    (mirrors TP/utils/common_config.py:34-41,64-90)."""
    import torch.nn as nn

    _activate("TaskPrompter")
    from easydict import EasyDict
    from models.transformers.taskprompter_swin import TaskPrompterSwin
    from models.transformers.taskprompter import ConvHead, DEConvHead
    from models.taskprompter_wrapper import TaskPrompterWrapper

    h, w = cfg["img_size"]
    E = cfg["embed_dim"]
    p = EasyDict(TASKS=EasyDict(NAMES=list(cfg["tasks"]), NUM_OUTPUT=dict(cfg["num_output"])),
                 prompt_len=cfg["prompt_len"], chan_embed_dim=cfg["chan_embed_dim"], chan_nheads=cfg["chan_nheads"],
                 img_ds_ratio=cfg["img_ds_ratio"], level_embed_dim=cfg["level_embed_dim"], final_embed_dim=cfg["f"],
                 backbone_channels=[2 * E, 4 * E, 8 * E, 8 * E],                      # common_config.py:36
                 ori_spatial_dim=[[h // st, w // st] for st in (8, 16, 32, 32)])       # :37-39
    if "dd_label_map_size" in cfg:
        p.dd_label_map_size = list(cfg["dd_label_map_size"])
    backbone = TaskPrompterSwin(p=p, img_size=tuple(cfg["img_size"]), patch_size=cfg["patch"], embed_dim=E,
                                depths=tuple(cfg["depths"]), num_heads=tuple(cfg["heads"]),
                                window_size=cfg["window"], drop_path_rate=0.15)
    head_cls = DEConvHead if cfg.get("head", "conv") == "deconv" else ConvHead
    heads = nn.ModuleDict({t: head_cls(cfg["f"], cfg["num_output"][t]) for t in cfg["tasks"]})
    return TaskPrompterWrapper(p, backbone, heads)


def build_invpt(cfg):
    """Reference TransformerNet (InvPT) for a config dict from oracle.configs.invpt()
    (mirrors IP/utils/common_config.py:15-21,39-51)."""
    import torch.nn as nn

    _activate("InvPT")
    from easydict import EasyDict
    from models.transformers.vit import VisionTransformer
    from models.transformers.transformer_decoder import MLPHead
    from models.transformer_net import TransformerNet

    H, W = cfg["img_size"]
    gh, gw = H // cfg["patch"], W // cfg["patch"]
    p = EasyDict(TASKS=EasyDict(NAMES=list(cfg["tasks"]), NUM_OUTPUT=dict(cfg["num_output"])),
                 embed_dim=cfg["embed_dim"], PRED_OUT_NUM_CONSTANT=cfg["pred_const"],
                 mtt_resolution_downsample_rate=cfg["down"])
    p.backbone_channels = [cfg["C"]] * 4
    p.spatial_dim = [[gh, gw]] * 4
    p.final_embed_dim = cfg["embed_dim"] + cfg["pred_const"]
    backbone = VisionTransformer(select_list=list(cfg["select"]), img_size=(H, W), patch_size=cfg["patch"],
                                 embed_dim=cfg["C"], depth=cfg["depth"], num_heads=cfg["heads"],
                                 drop_path_rate=0.15)
    heads = nn.ModuleDict({t: MLPHead(p.final_embed_dim, cfg["num_output"][t]) for t in cfg["tasks"]})
    return TransformerNet(p, backbone, p.backbone_channels, heads)
