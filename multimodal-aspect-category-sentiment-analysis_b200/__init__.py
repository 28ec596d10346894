"""fcmf_b200 -- B200-native (sm_100a) implementation of the FCMF fine-grained cross-modal fusion hot path.

    from importlib import import_module
    pkg = import_module("multimodal-aspect-category-sentiment-analysis_b200")    # or: import fcmf_b200
    model = pkg.FCMF(pretrained_path, num_labels=4, num_imgs=7, num_roi=4)

Everything numeric runs in ``libfcmf_b200.so`` (hand-written CUDA, C ABI in include/fcmf_b200.h); importing the
package does not require a GPU, calling the fusion path does.
"""
from . import synth                                                  # noqa: F401  (numpy/torch only)
from ._lib import LIB_PATH, exported_symbols, load                    # noqa: F401
from ._build import build                                            # noqa: F401
from .fcmf_framework import FCMF, FCMFEncoder, FCMFSeq2Seq           # noqa: F401
from . import ops, functional, fusion                                # noqa: F401

__all__ = ["FCMF", "FCMFEncoder", "FCMFSeq2Seq", "build", "load", "ops", "functional", "fusion", "synth"]
