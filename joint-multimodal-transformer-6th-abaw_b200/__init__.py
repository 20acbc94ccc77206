"""jmt_b200: B200-native (sm_100a) engine for the Joint Multimodal Transformer hot path.

Drop-in replacements (same names / constructor & forward signatures / state_dict keys) for the
reference modules named in SURVEY.md section 8b, backed by hand-written CUDA kernels behind the
C-ABI in include/jmt_b200.h.  CUDA only; no CPU or ATen fallback.
"""
from ._lib import LIB_PATH, launch_count, lib  # noqa: F401
from .modules import (FcLayer, JMTPipeline, FeatureConcatFC, Intra_modal_transformer_fusion,  # noqa: F401
                      MultimodalTransformer_w_JR, MultimodalTransformer_wo_JR, SingleBackbonePretrainer,
                      TemporalBlock, TemporalConvNet, TransformerEncoderBlock, TransformerEncoderLayer,
                      Two_transformers)
from .losses import CCCLoss, CCCLossMasked, LiveCCCLoss  # noqa: F401
from . import cccmetric  # noqa: F401
from . import padseq  # noqa: F401
from . import dist  # noqa: F401
from .graphs import GraphedStep  # noqa: F401
from . import valpost  # noqa: F401
from . import checkpoint  # noqa: F401
from . import features  # noqa: F401
