"""Drop-in mirror of the reference's ``fcmf_framework`` package for the fusion path: same module names, class
names, constructor/forward signatures and state_dict keys; the arithmetic runs in the
sm_100a kernels of ``libfcmf_b200.so``."""
from . import mm_modeling, roi_modeling, fcmf_pretraining, fcmf_multimodal  # noqa: F401
from .fcmf_pretraining import FCMFEncoder, FCMFSeq2Seq                       # noqa: F401
from .fcmf_multimodal import FCMF                                            # noqa: F401
