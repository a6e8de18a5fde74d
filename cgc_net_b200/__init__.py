"""cgc_net_b200 -- B200-native (sm_100a) graph-convolution hot path of CGC-Net.

Public surface (mirrors ``/root/reference/model/network.py``):

    from cgc_net_b200 import network
    model = network.SoftPoolingGcnEncoder(max_num_nodes, input_dim, hidden, out, True, True,
                                          hidden, num_classes, assign_ratio, [50], ...)

The arithmetic runs in ``csrc/libcgc_sm100.so`` (hand-written CUDA behind the C-ABI declared
in ``include/cgc_sm100.h``); importing the package does not need a GPU, calling it does.
"""
from . import graph, network, ops            # noqa: F401
from .network import SoftPoolingGcnEncoder, GNN_Module, DenseSAGEConv   # noqa: F401
from .graph import Data, Batch               # noqa: F401

__version__ = '0.1.0'
