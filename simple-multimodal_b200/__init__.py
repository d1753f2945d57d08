"""simple-multimodal_b200: B200-native (sm_100a) fusion heads for nl1xx/simple-multimodal.

Import with `importlib.import_module("simple-multimodal_b200")` or through the identifier-safe alias
module `simple_multimodal_b200` at the repository root.  `fusion_layers` is the drop-in for the
reference's `models.fusion_layers`.  Everything computes in libb200fusion.so;
importing works without a GPU, running does not (no fallback).
"""
from . import _lib, kernels, ops, mult_engine, fusion_layers            # noqa: F401
from ._lib import B200FusionError, LIB_PATH                              # noqa: F401
from .fusion_layers import (AdaptiveFusion, ContrastiveFusion, CrossModalTransformer, EarlyFusion, GraphFusion,   # noqa: F401
                            HierarchicalFusion, LateFusion, ModalityDropout, MultimodalTransformer)
from .ops import GradBucket, allreduce_gradients, global_batch_scale, manual_seed   # noqa: F401
from .staging import FeaturePrefetcher, bind_host_to_gpu                 # noqa: F401
from .graphs import GraphedTrainStep                                     # noqa: F401
from .optim import FusedAdamW                                            # noqa: F401
from .sequence_features import SequenceProjector, text_pooling_of        # noqa: F401
from . import prediction_heads                                           # noqa: F401
from .prediction_heads import AuxiliaryHeads, EmotionClassifier, SmoothedCrossEntropy   # noqa: F401

__all__ = ["EarlyFusion", "LateFusion", "MultimodalTransformer", "CrossModalTransformer", "GraphFusion", "ContrastiveFusion",
           "AdaptiveFusion", "HierarchicalFusion", "ModalityDropout", "allreduce_gradients", "GradBucket", "global_batch_scale", "manual_seed", "FeaturePrefetcher", "bind_host_to_gpu", "GraphedTrainStep", "FusedAdamW", "SequenceProjector", "text_pooling_of", "EmotionClassifier", "AuxiliaryHeads", "SmoothedCrossEntropy", "B200FusionError"]
