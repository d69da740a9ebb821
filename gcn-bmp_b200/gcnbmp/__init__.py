"""gcnbmp -- B200-native drop-in for the GCN-BMP message-passing hot path.

Host-side mirror of the reference's Chainer links (links.py) over a ctypes C-ABI
(_capi.py, include/gcnbmp.h) into hand-written sm_100a CUDA kernels (csrc/).
Importing this package loads libgcnbmp.so and fails loudly when it is missing.
"""
from . import _capi
from ._capi import BmpError, launch_count, reset_launch_count, MODE_F32, MODE_BF16
from . import functional
from .functional import pack_adjacency, unpack_adjacency
from . import metrics
from . import evaluate
from . import fused
from .links import (MAX_ATOMIC_NUM, functions, Link, ChainList, GraphLinear, GGNNUpdate, RelGCNUpdate,
                    GGNNReadout, GGNN, GGNNMono, RelGCN, GIN, GINUpdate, NFP, NFPUpdate, NFPReadout, BiMPM, NieFineCoattention, VQAParallelCoattention,
                    PoolingFineCoattention, AlternatingCoattention, ParallelCoattention, CircularParallelCoattention, GlobalCoattention, NeuralCoattention, FourierFineCoattention, DeepNieFineCoattention, VeryDeepNieFineCoattention, ExtremeDeepNieFineCoattention, HolE, HOLE, MLP, SymMLP, NTN, DistMult, BilinearDiag, GraphConvPredictorForPair,
                    sigmoid_cross_entropy, seed)

__version__ = "0.1.0"
