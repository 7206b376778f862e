"""CPU restatement of the reference's training losses -- TEST INFRASTRUCTURE (N3 "loss
reductions", and the scalar that N1's backward differentiates).

Per task (TaskPrompter/utils/common_config.py:211-237 picks them; InvPT is identical):
  semseg, human_parts  cross entropy with ignore regions           losses/loss_functions.py:16-60
  sal                  the same, class-balanced (binary)           :32-41
  edge                 balanced binary cross entropy, HED weights  :62-93
  normals              L1 on L2-normalised predictions             :144-176 (L1Loss(normalize=True))
  depth                L1 with an invalid-area mask (label == -1)  :144-176 (L1Loss(ignore_index=-1))
and their weighted sum, `MultiTaskLoss.forward` (losses/loss_schemes.py:26-39). Every function is a pure function of
(prediction, label) so autograd through it and through oracle.taskprompter_ref.forward gives the reference's
parameter gradients (pinned in tests/test_oracle.py::test_loss_* against tests/golden/losses.pt, which
oracle/make_golden.py produces with the reference's own loss modules and autograd through the reference model)."""
import torch
import torch.nn.functional as F

IGNORE = 255


def cross_entropy(pred, label, ignore_index=IGNORE, balanced=False):
    """pred [B,C,H,W] logits, label [B,1,H,W] (any dtype holding class ids / ignore_index)."""
    tgt = label[:, 0].long()
    keep = tgt != ignore_index
    w = None
    if balanced:                                      # :32-41 -- binary: weight (1 - w_pos, w_pos)
        kept = tgt[keep].to(pred.dtype)
        w_pos = (1.0 - kept).sum() / kept.numel()
        w = torch.stack((1.0 - w_pos, w_pos))
    logp = F.log_softmax(pred, dim=1)
    safe = tgt.clamp(0, pred.shape[1] - 1)
    nll = -logp.gather(1, safe[:, None])[:, 0]
    if w is not None:
        nll = nll * w[safe]
    total = (nll * keep).sum()                        # reduction='none' then sum; ignored pixels contribute 0
    return total / max(int(keep.sum()), 1)            # :53-55 (divides by the number of valid pixels, unweighted)


def balanced_bce(pred, label, pos_weight=None, ignore_index=IGNORE):
    """pred, label [B,1,H,W]; HED-style weighting unless pos_weight is given (:62-93)."""
    keep = label != ignore_index
    y = label[keep]
    x = pred[keep]
    if pos_weight is None:
        w = (1.0 - y).sum() / y.numel()
        if w == 1.0:
            return pred.sum() * 0.0
    else:
        w = torch.as_tensor(pos_weight, dtype=pred.dtype)
    # BCEWithLogits(pos_weight = w / (1 - w)) * (1 - w)  ==  -(w y log s(x) + (1 - w)(1 - y) log(1 - s(x)))
    per = w * y * F.softplus(-x) + (1.0 - w) * (1.0 - y) * F.softplus(x)
    return per.mean()


def l1(pred, label, normalize=False, ignore_index=0, ignore_invalid_area=True):
    """L1Loss (:144-176): mean absolute error over pixels whose label differs from ignore_index in EVERY channel;
    the denominator counts PIXELS (not pixels x channels)."""
    if normalize:
        pred = F.normalize(pred, p=2, dim=1)
    if ignore_invalid_area:
        keep = (label != ignore_index).all(dim=1, keepdim=True)
    else:
        keep = torch.ones_like(label[:, :1], dtype=torch.bool)
    diff = (pred - label).abs() * keep
    return diff.sum() / max(int(keep.sum()), 1)


def task_loss(task, pred, label, edge_w=0.95, ignore_invalid_area_depth=True):
    if task in ("semseg", "human_parts"):
        return cross_entropy(pred, label)
    if task == "sal":
        return cross_entropy(pred, label, balanced=True)
    if task == "edge":
        return balanced_bce(pred, label, pos_weight=edge_w)
    if task == "normals":
        return l1(pred, label, normalize=True, ignore_index=IGNORE)
    if task == "depth":
        return l1(pred, label, ignore_index=-1, ignore_invalid_area=ignore_invalid_area_depth)
    raise ValueError(task)


def multi_task_loss(preds, labels, tasks, weights, **kw):
    """MultiTaskLoss.forward (loss_schemes.py:26-39): {task: loss, 'total': sum_t weights[t] * loss_t}."""
    out = {t: task_loss(t, preds[t], labels[t], **kw) for t in tasks}
    out["total"] = torch.stack([weights[t] * out[t] for t in tasks]).sum()
    return out
