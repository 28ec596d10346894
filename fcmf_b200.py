"""Import alias: ``import fcmf_b200`` == the package ``multimodal-aspect-category-sentiment-analysis_b200``
(whose directory name is not a valid Python identifier)."""
import importlib
import sys

_pkg = importlib.import_module("multimodal-aspect-category-sentiment-analysis_b200")
sys.modules[__name__] = _pkg
