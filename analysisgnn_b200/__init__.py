"""analysisgnn_b200: B200-native kernels and drop-in modules for AnalysisGNN's heterogeneous
message-passing hot path.  CUDA only -- there is no CPU path."""
import os as _os

import torch as _torch

__version__ = "0.1.0"


def set_fp32_exact(enabled: bool = True) -> None:
    """The parity contract is the reference's fp32 arithmetic (1e-5 relative).  PyTorch lets
    cuDNN run fp32 RNNs / convolutions in TF32 by default (``torch.backends.cudnn.allow_tf32``),
    which is ~1e-3 away; the library GEMMs / cuDNN GRU still on the path must therefore run
    with TF32 off, in forward and backward.  Set ``AGNN_ALLOW_TF32=1`` to leave the flags alone."""
    _torch.backends.cudnn.allow_tf32 = not enabled
    _torch.backends.cuda.matmul.allow_tf32 = not enabled


if _os.environ.get("AGNN_ALLOW_TF32", "0") != "1":
    set_fp32_exact(True)
