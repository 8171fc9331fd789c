"""src/models/frame_transformer.py of the reference, B200-native (see tvt_b200.hostapi.frame_transformer)."""
from tvt_b200.hostapi.frame_transformer import (FrameTransformer, ImgResNet, PositionalEncoding, TransformerBase,  # noqa: F401
                                                VidResNet)
