"""One non-graph TaskPrompter cfg4 forward bracketed by cudaProfilerStart/Stop, for
`ncu --profile-from-start off ...` (launch list and --set full captures). Not a benchmark."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mtt_b200
from mtt_b200 import configs

cfg_name = sys.argv[1] if len(sys.argv) > 1 else "tp_cfg4"
mode = sys.argv[2] if len(sys.argv) > 2 else "parity"
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 4
if cfg_name.startswith("tps_"):
    from mtt_b200 import taskprompter_swin as TP
    cfg = configs.taskprompter_swin(cfg_name)
elif cfg_name.startswith("tp_"):
    from mtt_b200 import taskprompter as TP
    cfg = configs.taskprompter(cfg_name)
else:
    from mtt_b200 import invpt as TP
    cfg = configs.invpt(cfg_name)
dev = torch.device("cuda:0")
torch.manual_seed(0)
with torch.device(dev):
    model = TP.build_from_config(cfg, nsplit=2 if mode == "parity" else 1, use_graph=False).eval()
x = torch.randn(batch, 3, *cfg["img_size"], device=dev)
with torch.no_grad():
    model(x)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    model(x)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("done")
