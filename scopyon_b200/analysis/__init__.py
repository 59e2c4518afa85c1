"""``scopyon.analysis`` on the GPU: blob and spot detection
(``/root/reference/src/scopyon/analysis/__init__.py:1-4``).  The hidden-Markov trajectory
models of the reference (``analysis/hmm.py``) are CPU post-processing outside the image path
and are not provided."""
from .spot_detection import *

__all__ = ["blob_detection", "spot_detection"]
