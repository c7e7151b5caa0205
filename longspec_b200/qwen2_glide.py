"""Drop-in for the reference's `qwen2_glide` module (longspec/test/qwen2_glide.py).

Same algorithm as LlamaGlide; the twin differs by q/k/v biases in the target (qwen2.py:278-281),
`output_ids` pre-filled with 0 and termination on the `eos_id` argument (qwen2_glide.py:766,949).
"""
from __future__ import annotations

from .llama_glide import LlamaGlide


class Qwen2Glide(LlamaGlide):
    variant = "qwen"
