"""``models`` package with the reference's module layout (src/models/*.py) so that the reference's own entry script and
callers resolve UNCHANGED against the B200-native classes: run with this ``src/`` directory on ``sys.path`` (the
reference's ``python main.py`` is started inside ``src/``) and

    from models.transformer import SimpleTransformer          # src/main.py:15
    from models.frame_transformer import FrameTransformer     # src/main.py:16

bind to ``tvt_b200.hostapi`` (sm_100a kernels behind the same constructors / hooks / state_dict keys).  Only the hot-path
model files are mirrored (transformer, frame_transformer, TPN, vit, collabgating); LSTM / contrastive baselines and the CNN
feature extractors are out of scope."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
import tvt_b200  # noqa: E402,F401  (registers the package alias)
