"""src/models/collabgating.py of the reference, B200-native (see tvt_b200.hostapi.collabgating)."""
from tvt_b200.hostapi.collabgating import CollaborativeGating  # noqa: F401
