"""CPU oracle for the LongSpec draft-then-tree-verify decode loop.

THIS PACKAGE IS TEST INFRASTRUCTURE.  It is a CPU restatement (torch-CPU fp32 for
the floating-point parts, numpy integers for the tree bookkeeping) of the
reference algorithm, function by function, each citing the reference file:line
it follows (paths relative to /root/reference/).  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import it - as the checker or the timed CPU baseline, never as the
product path.  `longspec_b200/` never imports it.

Pinning: the reference ships no tests and no golden
vectors.  The oracle is pinned instead against outputs of the reference's own
code run in the build container through `tests/golden/ref_shim.py`
(`tests/golden/gen_golden.py` writes `tests/golden/*.npz`; `tests/test_oracle_*`
re-checks the oracle against those fixtures on every run), and against the
reference's structural invariant tree(T=0) == vanilla(T=0).

Third-party arithmetic that is not in /root/reference: flash_attn==2.6.3
(`flash_attn_func`, `flash_attn_with_kvcache`; longspec/test/requirements.txt:2)
and triton==3.1.0 (compiles triton_tree_attn.py).  `oracle/flash_ref.py` restates
flash-attn's published semantics; the reference's Triton kernel itself was run
under TRITON_INTERPRET=1 to pin `oracle.attention.tree_mask_attention`
(fixture `tests/golden/triton_tree_attn.npz`).
"""
