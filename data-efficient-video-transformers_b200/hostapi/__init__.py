"""Host-side mirror of the reference's model API (src/models/*.py): same class names, constructor
arguments, forward signatures, Lightning hook names and state_dict keys; the arithmetic underneath runs
on the package's sm_100a kernels through torch.autograd.Functions (see ..functions)."""
from .transformer import PositionalEncoding, SimpleTransformer  # noqa: F401
from .frame_transformer import TransformerBase, FrameStream, FrameTransformer, ImgResNet, VidResNet  # noqa: F401
from .TPN import (Reasoning, sum_group, SpatialPyramid, Feature_Pyramid_low, Feature_Pyramid_Mid,  # noqa: F401
                  Feature_Pyramid_High, TPN, ResNetMaps)
from .fusion import CrossModalBlock, ExpertStream, FusionTransformer, DistillationTrainer  # noqa: F401
from .inference import EvalBuffer, GraphedForward, GraphedTrainStep, REFERENCE_THRESHOLDS  # noqa: F401
from .collabgating import CollaborativeGating  # noqa: F401
from .loader import FeatureAugment  # noqa: F401
from . import vit  # noqa: F401,E402
