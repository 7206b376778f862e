"""One training step (mtt_b200.train.TrainStep.step) bracketed by cudaProfilerStart/Stop, for
`ncu --profile-from-start off --metrics gpu__time_duration.sum --csv ...` (launch list). Not a benchmark."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mtt_b200
from mtt_b200 import configs, taskprompter as TP
from mtt_b200.train import TrainStep
import bench

cfg_name = sys.argv[1] if len(sys.argv) > 1 else "tp_cfg4"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 4
cfg = configs.taskprompter(cfg_name)
dev = torch.device("cuda:0")
torch.manual_seed(0)
with torch.device(dev):
    model = TP.build_from_config(cfg, use_graph=False)
ts = TrainStep(model)
crit, _ = bench._train_criterion(cfg)
g = torch.Generator().manual_seed(1)
x = torch.randn(batch, 3, *cfg["img_size"], generator=g).to(dev)
y = {t: v.to(dev) for t, v in bench._train_labels(cfg, batch, g).items()}
with torch.no_grad():
    ts.step(x, y, crit)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    ts.step(x, y, crit)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("done")
