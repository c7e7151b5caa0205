"""`from qwen2_glide import Qwen2Glide` for the unchanged reference driver (inference_qwq.py:1)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from longspec_b200.qwen2_glide import Qwen2Glide  # noqa: E402,F401
