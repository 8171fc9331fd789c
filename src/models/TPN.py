"""src/models/TPN.py of the reference, B200-native (see tvt_b200.hostapi.TPN)."""
from tvt_b200.hostapi.TPN import (TPN, Feature_Pyramid_High, Feature_Pyramid_Mid, Feature_Pyramid_low, Reasoning,  # noqa: F401
                                  sum_group)
