"""CPU restatement of the reference's prediction post-processing -- TEST INFRASTRUCTURE.

`get_output` (TaskPrompter/utils/utils.py:27-63, identical in InvPT/utils/utils.py:18-48): per task, from
NCHW logits [B,C,H,W] to the map the evaluation / visualisation code consumes."""
import torch
import torch.nn.functional as F


def get_output(output, task):
    if task == "normals":                                   # utils.py:29-31
        o = output.permute(0, 2, 3, 1)
        return (F.normalize(o, p=2, dim=3) + 1.0) * 255 / 2.0
    if task in ("semseg", "human_parts"):                   # :33-43
        return output.permute(0, 2, 3, 1).max(dim=3)[1]
    if task == "edge":                                      # :45-47
        return torch.squeeze(255 * 1 / (1 + torch.exp(-output.permute(0, 2, 3, 1))), dim=3)
    if task == "sal":                                       # :49-51
        return F.softmax(output.permute(0, 2, 3, 1), dim=3)[:, :, :, 1] * 255
    if task == "depth":                                     # :53-55
        return output.clamp(min=0.).permute(0, 2, 3, 1)
    raise ValueError(task)
