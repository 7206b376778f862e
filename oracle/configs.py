"""Named model configurations shared by the oracle, the tests and the bench: re-exported from the package
(mtt_b200/configs.py holds the data; it contains no computation)."""
from mtt_b200.configs import *  # noqa: F401,F403
from mtt_b200.configs import invpt, taskprompter, taskprompter_swin  # noqa: F401
