"""beyond_deep_ensembles_b200 — B200-native posterior-update path behind the optimizer API of
Feuermagier/Beyond_Deep_Ensembles (src/algos).

Host code is Python/PyTorch (plumbing: device memory, streams, torch.distributed); the
arithmetic is hand-written sm_100a CUDA in lib/libbde_b200.so behind the C-ABI of
include/bde_b200.h.  There is no CPU or eager-PyTorch fallback: without the library the
package raises.
"""
from .algo import BayesianOptimizer, LastLayerBayesianOptimizer
from .bbb import BBBOptimizer, GaussianPrior, MixturePrior
from .ensemble import DeepEnsemble
from .install import install
from .sharded_closure import ColumnShardedModel
from .ivorn import iVONOptimizer
from .svgd import SVGDOptimizer, rbf
from .swag import SwagOptimizer
from .util import GaussianParameter

__all__ = [
    "BayesianOptimizer", "LastLayerBayesianOptimizer", "BBBOptimizer", "GaussianPrior", "MixturePrior",
    "DeepEnsemble", "install", "ColumnShardedModel", "iVONOptimizer", "SVGDOptimizer", "rbf", "SwagOptimizer", "GaussianParameter",
]
