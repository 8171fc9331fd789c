"""src/models/transformer.py of the reference, B200-native (see tvt_b200.hostapi.transformer)."""
from tvt_b200.hostapi.transformer import PositionalEncoding, SimpleTransformer  # noqa: F401
