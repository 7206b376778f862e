"""Graph-replay timing of any config (TaskPrompter tp_* or InvPT ip_*): python scripts/time_model.py ip_cfg3 4 [parity|speed]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mtt_b200
from mtt_b200 import configs

name = sys.argv[1]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
mode = sys.argv[3] if len(sys.argv) > 3 else "parity"
ns = 2 if mode == "parity" else 1
dev = torch.device("cuda:0")
torch.manual_seed(0)
if name.startswith("tp_"):
    from mtt_b200 import taskprompter as M
    cfg = configs.taskprompter(name)
else:
    from mtt_b200 import invpt as M
    cfg = configs.invpt(name)
with torch.device(dev):
    model = M.build_from_config(cfg, nsplit=ns, use_graph=True).eval()
x = torch.randn(B, 3, *cfg["img_size"], device=dev)
with torch.no_grad():
    for _ in range(3):
        model(x)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    s.record()
    for _ in range(n):
        model(x)
    e.record()
    torch.cuda.synchronize()
ms = s.elapsed_time(e) / n
print(f"{name} B={B} {mode}: {ms:.3f} ms/step  {B / ms * 1e3:.1f} img/s  peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
