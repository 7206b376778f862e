"""Generates tests/golden/*.pt by running the UNMODIFIED reference (imported from /root/reference through
oracle/shim) in the build container. TEST INFRASTRUCTURE.

Weights come from the deterministic oracle initialisers (oracle.taskprompter_ref.init_state_dict,
oracle.invpt_ref.init_state_dict: fixed torch CPU generator seeds) loaded into the reference model
with strict=True, so a fixture only needs to carry the input, the reference's outputs and a checksum
of the parameters -- the GPU box regenerates identical weights without /root/reference.

    python -m oracle.make_golden            # rewrites tests/golden/
"""
import hashlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import configs, ref_loader  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def sd_checksum(sd):
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def make_taskprompter(name, seed, batch):
    from oracle import taskprompter_ref as R

    cfg = configs.taskprompter(name)
    sd = R.init_state_dict(cfg, seed=seed)
    model = ref_loader.build_taskprompter(cfg).eval()
    model.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(seed + 1000)
    x = torch.randn(batch, 3, *cfg["img_size"], generator=g)
    with torch.no_grad():
        y = model(x)
    return {"family": "taskprompter", "cfg": name, "seed": seed, "x": x, "out": {k: v.clone() for k, v in y.items()},
            "sd_sha256": sd_checksum(sd), "torch": torch.__version__,
            "made_by": "oracle/make_golden.py from the unmodified reference forward (eval, fp32, CPU)"}


def make_invpt(name, seed, batch):
    from oracle import invpt_ref as R

    cfg = configs.invpt(name)
    sd = R.init_state_dict(cfg, seed=seed)
    model = ref_loader.build_invpt(cfg).eval()
    model.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(seed + 1000)
    x = torch.randn(batch, 3, *cfg["img_size"], generator=g)
    with torch.no_grad():
        y = model(x)
    out = {k: v.clone() for k, v in y.items() if k != "inter_preds"}
    # inter_preds kept only for the tiny config (fixture size)
    inter = {k: v.clone() for k, v in y["inter_preds"].items()} if name == "ip_tiny" else None
    return {"family": "invpt", "cfg": name, "seed": seed, "x": x, "out": out, "inter_preds": inter,
            "sd_sha256": sd_checksum(sd), "torch": torch.__version__,
            "made_by": "oracle/make_golden.py from the unmodified reference forward (eval, fp32, CPU)"}


def make_preproc():
    """Golden vectors for the inference pre-processing: the reference's OWN Normalize / ToTensor classes
    (TaskPrompter/data/transforms.py) around cv2.cvtColor + cv2.resize exactly as TaskPrompter/inference.py:66-81,
    :93-115, :127-133 chains them (inference.py itself parses argv and loads a checkpoint at import, so its three-line
    DirectResize is restated here around the same cv2 call)."""
    import cv2
    import numpy as np

    ref_loader._activate("TaskPrompter")
    from data import transforms as T

    rng = np.random.default_rng(7)
    cases = []
    for (h, w, H, W) in [(37, 53, 64, 48), (96, 120, 64, 80), (50, 70, 50, 70)]:
        bgr = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        img = cv2.cvtColor(bgr.astype(np.float32), cv2.COLOR_BGR2RGB)          # inference.py:127-128
        sample = T.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])({"image": img})
        sample["image"] = cv2.resize(sample["image"], (W, H), interpolation=cv2.INTER_LINEAR)   # DirectResize
        out = T.ToTensor()(sample)["image"].unsqueeze(0)
        cases.append({"bgr_u8": torch.from_numpy(bgr), "out_hw": (H, W), "out": out.clone()})
    return {"family": "preproc", "cases": cases, "cv2": cv2.__version__,
            "made_by": "oracle/make_golden.py: reference Normalize/ToTensor + cv2.resize (TP/inference.py pipeline)"}


def main():
    if not ref_loader.available():
        raise SystemExit("reference not found (set MTT_REFERENCE or mount /root/reference)")
    os.makedirs(GOLD, exist_ok=True)
    jobs = [("taskprompter", "tp_tiny", 3, 2), ("taskprompter", "tp_tiny1", 4, 2), ("taskprompter", "tp_tiny_de", 8, 2)]
    if os.environ.get("MTT_GOLDEN_ONLY") in ("tp_tiny_de",):
        jobs = [j for j in jobs if j[1] == os.environ["MTT_GOLDEN_ONLY"]]
    elif os.path.exists(os.path.join(ROOT, "oracle", "invpt_ref.py")):
        jobs += [("invpt", "ip_tiny", 5, 2), ("invpt", "ip_cfg1", 6, 2)]
    path = os.path.join(GOLD, "preproc.pt")
    torch.save(make_preproc(), path)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")
    if os.environ.get("MTT_GOLDEN_ONLY") == "preproc":
        return
    for fam, name, seed, batch in jobs:
        fx = make_taskprompter(name, seed, batch) if fam == "taskprompter" else make_invpt(name, seed, batch)
        path = os.path.join(GOLD, f"{name}.pt")
        torch.save(fx, path)
        print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


if __name__ == "__main__":
    main()
