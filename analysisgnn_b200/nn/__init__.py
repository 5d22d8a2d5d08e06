"""Drop-in ``torch.nn.Module``s for the message-passing hot path.

Same class names, constructor arguments, forward arguments and ``state_dict``
keys as the modules they replace, so reference checkpoints load unchanged:

* ``intree``  -- ``SageConvScatter``, ``HeteroConv``, ``MetricalConvLayer``,
  ``MetricalGNN`` of ``analysisgnn/models/core/{gnn,hgnn}.py``;
* ``hetero``  -- the PyG / graphmuse shaped encoders built at
  ``analysisgnn/models/analysis.py:444-473`` (``HybridGNN``, ``HybridHGT``,
  ``MetricalGNN``) and their layers;
* ``shell``   -- the hot-path part of ``TorchAnalysisGNN``
  (``analysisgnn/models/analysis.py:421-591``);
* ``layers``  -- ``Linear`` / ``LayerNorm`` / ``GRU`` with ``torch.nn``'s parameters on libagnn's kernels.

Every forward runs on libagnn.so's CUDA kernels; there is no CPU path.
"""
from .intree import (GATConvLayer, HeteroConv, MetricalConvLayer, MetricalGNN, OnsetEmbedding,  # noqa: F401
                     RelEdgeConv, ResGatedGraphConv, SageConvScatter)
from .hetero import (HeteroSAGELayer, HeteroSAGEStack, HGTConv, HeteroHGTStack, HybridGNN, HybridHGT,  # noqa: F401
                     SAGEConv, SequenceBranch)
from .layers import GRU, LayerNorm, Linear  # noqa: F401
from .shell import AnalysisEncoder, CrossEntropyLoss, MultiTaskLoss, multitask_ce, onset_pool  # noqa: F401
