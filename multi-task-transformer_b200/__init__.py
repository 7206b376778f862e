"""mtt_b200 -- B200-native (sm_100a) TaskPrompter / InvPT forward hot path.

Host side mirrors the reference's nn.Module boundaries
(TaskPrompter/models/transformers/taskprompter.py, TaskPrompter/models/taskprompter_wrapper.py,
InvPT/models/transformers/{vit,transformer_decoder,invpt}.py, InvPT/models/transformer_net.py) and
calls hand-written CUDA kernels through the C ABI in include/mtt_b200.h. No CPU / eager fallback.
"""
from . import lib  # noqa: F401

__all__ = ["lib"]
