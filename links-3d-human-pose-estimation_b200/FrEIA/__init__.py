"""FrEIA-compatible shim (github.com/VLL-HD/FrEIA) for the two classes LInKs uses:
``FrEIA.framework.SequenceINN`` and ``FrEIA.modules.AllInOneBlock`` (reference call sites
train_leg_torso_lifter.py:352-363, train_full_pose_norm_flow.py:47-49).  Same constructor arguments,
call convention ``inn(x, rev=False) -> (out, log_jac_det)`` and state-dict keys; the arithmetic runs on the
fused sm_100a flow kernels of links_b200."""
from . import framework, modules  # noqa: F401
