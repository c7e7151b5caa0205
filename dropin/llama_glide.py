"""`from llama_glide import LlamaGlide` for the unchanged reference drivers
(inference_long-bench.py:1): put this directory on PYTHONPATH."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from longspec_b200.llama_glide import LlamaGlide  # noqa: E402,F401
