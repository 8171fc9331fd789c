"""src/models/vit.py of the reference, B200-native (see tvt_b200.hostapi.vit)."""
from tvt_b200.hostapi.vit import Attention, FeedForward, PreNorm, Transformer, ViViT  # noqa: F401
