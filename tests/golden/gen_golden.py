"""Generate the golden fixtures under tests/golden/ by RUNNING THE REFERENCE'S OWN CODE.

Run in the build container only (needs /root/reference):

    python tests/golden/gen_golden.py            # everything except the Triton fixture
    TRITON_INTERPRET=1 python tests/golden/gen_golden.py --triton

Every array written here is an input to, or an output of, unmodified reference code
(imported through tests/golden/ref_shim.py).  The oracle (oracle/) and the CUDA path
are then checked against these files; the files travel to the GPU box, the
reference does not.
"""
from __future__ import annotations

import argparse
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_shim  # noqa: E402
from longspec_b200.synthetic import PRESETS, synthetic_state_dict, synthetic_prompt  # noqa: E402


def rng(seed):
    g = torch.Generator()
    g.manual_seed(seed)
    return g


# ---------------------------------------------------------------------------
def random_tree(shape, g):
    """Grow a random beam-style tree mask exactly like llama_glide.py:1022,1069-1071."""
    offs = [1]
    for c in shape:
        offs.append(offs[-1] + c)
    f = offs[-1]
    m = torch.zeros(f, f, dtype=torch.int64)
    m[:, 0] = 1
    for r in range(1, offs[1]):
        m[r, r] += 1
    for lv in range(1, len(shape)):
        lo, hi, nxt = offs[lv - 1], offs[lv], offs[lv + 1]
        parents = torch.randint(lo, hi, (nxt - hi,), generator=g)
        for r, p in enumerate(parents.tolist()):
            m[hi + r] = m[p]
            m[hi + r, hi + r] += 1
    return m, offs


def gen_tree_verification(out):
    """Reference LlamaGlide.tree_verification (llama_glide.py:1128-1175) on random trees."""
    ref_shim.install()
    import llama_glide as ref_lg
    cases = []
    g = rng(11)
    shapes = [[4, 4], [2, 2], [4, 16, 16, 16, 16], [1, 1, 1, 1, 1], [8, 32, 32, 32, 32], [3, 5, 2]]
    Hk, D = 2, 16
    for ci, shape in enumerate(shapes * 3):
        mask, offs = random_tree(shape, g)
        f = offs[-1]
        V = 50
        all_spec = torch.randint(0, V, (1, f), generator=g)
        llm_pred = torch.randint(0, V, (1, f), generator=g)
        # plant a verified chain of random depth so that deep acceptance is exercised
        depth = int(torch.randint(0, len(shape) + 1, (1,), generator=g))
        node = 0
        for _ in range(depth):
            kids = [u for u in range(1, f) if mask[u].sum() == mask[node].sum() + 1 and
                    all(mask[u, j] >= mask[node, j] for j in range(f)) and mask[u, node] == 1]
            if not kids:
                break
            kid = kids[int(torch.randint(0, len(kids), (1,), generator=g))]
            llm_pred[0, node] = all_spec[0, kid]
            node = kid
        cache_len = int(torch.randint(5, 40, (1,), generator=g))
        S_max = cache_len + f + 8
        K = torch.randn(1, S_max, Hk, D, generator=g)
        Vc = torch.randn(1, S_max, Hk, D, generator=g)
        # stub "self" carrying exactly what the method reads
        attn = types.SimpleNamespace(K_Cache=K.clone(), V_Cache=Vc.clone(), num_key_value_heads=Hk, head_dim=D)
        stub = types.SimpleNamespace(
            range_tensor=torch.arange(0, 1024)[None, :],
            reverse_range_tensor=torch.arange(-1024, -32 + 1).unsqueeze(0),
            oned_range_tensor=torch.arange(0, 1024),
            diag_matrix=torch.eye(1024, dtype=torch.int64)[None],
            model=types.SimpleNamespace(layers=[types.SimpleNamespace(self_attn=attn)]))
        fn = ref_lg.LlamaGlide.tree_verification
        fn = getattr(fn, "__wrapped__", fn)
        acc_ids, acc_num, dbl = fn(stub, all_spec, llm_pred, mask[None], torch.tensor([cache_len]),
                                   non_leaf_len=offs[-2])
        cases.append(dict(shape=np.array(shape), all_spec=all_spec[0].numpy(), llm_pred=llm_pred[0].numpy(),
                          tree_mask=mask.numpy(), cache_len=np.int64(cache_len), non_leaf_len=np.int64(offs[-2]),
                          acc_ids=acc_ids[0].numpy(), acc_num=acc_num.numpy(), double_input=dbl.numpy(),
                          K_before=K[0, cache_len:cache_len + f].numpy().astype(np.float16),
                          K_after=attn.K_Cache[0, cache_len:cache_len + f].numpy().astype(np.float16),
                          V_before=Vc[0, cache_len:cache_len + f].numpy().astype(np.float16),
                          V_after=attn.V_Cache[0, cache_len:cache_len + f].numpy().astype(np.float16)))
    flat = {"n": np.int64(len(cases))}
    for i, c in enumerate(cases):
        for k, v in c.items():
            flat[f"c{i}_{k}"] = v
    np.savez_compressed(os.path.join(out, "tree_verification.npz"), **flat)
    print("tree_verification:", len(cases), "cases")


def gen_target_tree_part(out):
    """Reference LlamaAttention.tree_part_fwd (llama.py:394-421) called directly."""
    ref_shim.install()
    import llama as ref_llama
    g = rng(21)
    flat = {}
    cases = [(5, 2, 1, 7, False), (11, 4, 2, 20, True), (11, 2, 2, 3, False)]
    for ci, (S, H, Hk, n0, last) in enumerate(cases):
        D = 128
        mask, _ = random_tree([2, 2] if S == 5 else [4, 4], g)
        # verify-style mask: acc tokens + tree + pad, use plain random lower-tri for variety
        vm = torch.tril(torch.ones(S, S, dtype=torch.int64))
        f = mask.shape[0]
        a = 1
        vm[a:a + f - 1, a:a + f - 1] = mask[1:, 1:]
        vm = torch.tril(vm)
        q = torch.randn(1, S, H, D, generator=g)
        k = torch.randn(1, S, Hk, D, generator=g)
        v = torch.randn(1, S, Hk, D, generator=g)
        prefix_lse = torch.randn(1, H, S, generator=g) * 2 + 1
        K = torch.zeros(1, n0 + S + 4, Hk, D)
        Vc = torch.zeros(1, n0 + S + 4, Hk, D)
        stub = types.SimpleNamespace(range_indices=torch.arange(1024), K_Cache=K, V_Cache=Vc, num_heads=H,
                                     num_key_value_heads=Hk, last_layer=last, softmax_scale=1 / (128 ** 0.5))
        fn = ref_llama.LlamaAttention.tree_part_fwd
        fn = getattr(fn, "__wrapped__", fn)
        cur, w = fn(stub, q, k, v, vm[None], torch.tensor([n0]), prefix_lse, 1, S)
        flat.update({f"c{ci}_q": q.numpy(), f"c{ci}_k": k.numpy(), f"c{ci}_v": v.numpy(),
                     f"c{ci}_mask": vm.numpy(), f"c{ci}_prefix_lse": prefix_lse.numpy(),
                     f"c{ci}_n0": np.int64(n0), f"c{ci}_last": np.int64(last),
                     f"c{ci}_cur": cur.contiguous().numpy(), f"c{ci}_w": w.contiguous().numpy(),
                     f"c{ci}_K_after": K.numpy(), f"c{ci}_V_after": Vc.numpy()})
    flat["n"] = np.int64(len(cases))
    np.savez_compressed(os.path.join(out, "target_tree_part.npz"), **flat)
    print("target_tree_part:", len(cases), "cases")


def gen_draft_tree_part(out):
    """Reference GlideAttention.tree_part_fwd (llama_glide.py:331-359) called directly."""
    ref_shim.install()
    import llama_glide as ref_lg
    g = rng(31)
    flat = {}
    shape = [4, 4, 4]
    mask, offs = random_tree(shape, g)
    ci = 0
    for lv in (1, 2):
        lo, hi = offs[lv - 1], offs[lv]
        M, N = hi - lo, hi
        H, Hk, D, n0 = 4, 2, 128, 9
        cur_mask = mask[lo:hi, :hi]
        q = torch.randn(1, M, H, D, generator=g)
        k = torch.randn(1, M, Hk, D, generator=g)
        v = torch.randn(1, M, Hk, D, generator=g)
        K = torch.randn(1, n0 + N + 4, Hk, D, generator=g)
        Vc = torch.randn(1, n0 + N + 4, Hk, D, generator=g)
        K0, V0 = K.clone(), Vc.clone()
        prefix_lse = torch.randn(1, H, M, generator=g)
        stub = types.SimpleNamespace(range_indices=torch.arange(1024), K_Cache=K, V_Cache=Vc, num_heads=H,
                                     num_key_value_heads=Hk, num_key_value_groups=H // Hk,
                                     softmax_scale=1 / (128 ** 0.5))
        fn = ref_lg.GlideAttention.tree_part_fwd
        fn = getattr(fn, "__wrapped__", fn)
        cur, w = fn(stub, q, k, v, cur_mask[None], torch.tensor([n0]), prefix_lse, 1, M)
        flat.update({f"c{ci}_q": q.numpy(), f"c{ci}_k": k.numpy(), f"c{ci}_v": v.numpy(),
                     f"c{ci}_mask": cur_mask.numpy(), f"c{ci}_prefix_lse": prefix_lse.numpy(),
                     f"c{ci}_n0": np.int64(n0), f"c{ci}_K_before": K0.numpy(), f"c{ci}_V_before": V0.numpy(),
                     f"c{ci}_cur": cur.contiguous().numpy(), f"c{ci}_w": w.contiguous().numpy(),
                     f"c{ci}_K_after": K.numpy(), f"c{ci}_V_after": Vc.numpy()})
        ci += 1
    flat["n"] = np.int64(ci)
    np.savez_compressed(os.path.join(out, "draft_tree_part.npz"), **flat)
    print("draft_tree_part:", ci, "cases")


def gen_triton(out):
    """Reference Triton `_fwd_kernel` (triton_tree_attn.py:115-251) under the Triton
    interpreter, launched with the launcher's argument list (55-68) and the non-A100
    fallback tile config (111)."""
    assert os.environ.get("TRITON_INTERPRET") == "1", "run with TRITON_INTERPRET=1"
    ref_shim.install()
    import triton_tree_attn as tta
    g = rng(41)
    flat = {}
    ci = 0
    shape = [4, 16, 16]
    mask, offs = random_tree(shape, g)
    for (lo, hi) in [(1, 5), (5, 21)]:
        M, N = hi - lo, hi
        B, H, Hk, D = 1, 4, 2, 128
        q = torch.randn(B, H, M, D, generator=g)
        k = torch.randn(B, Hk, N, D, generator=g)
        v = torch.randn(B, Hk, N, D, generator=g)
        tm = mask[lo:hi, :hi][None].contiguous()
        o = torch.empty_like(q)
        L = torch.empty(B, H, M, dtype=torch.float32)
        BM = BN = 32
        grid = ((M + BM - 1) // BM, H, B)
        tta._fwd_kernel[grid](
            q, k, v, tm, 1.0 / (D ** 0.5), 0, 0, L, o,
            q.stride(0), q.stride(1), q.stride(2), q.stride(3),
            k.stride(0), k.stride(1), k.stride(2), k.stride(3),
            v.stride(0), v.stride(1), v.stride(2), v.stride(3),
            tm.stride(0), tm.stride(1), tm.stride(2),
            o.stride(0), o.stride(1), o.stride(2), o.stride(3),
            B, H, M, N, N - M, H // Hk,
            BLOCK_M=BM, BLOCK_N=BN, BLOCK_DMODEL=D, LARGER_M=M > N,
            DIVISIBLE_M=M % BM == 0, DIVISIBLE_N=N % BN == 0)
        flat.update({f"c{ci}_q": q.numpy(), f"c{ci}_k": k.numpy(), f"c{ci}_v": v.numpy(),
                     f"c{ci}_mask": tm[0].numpy(), f"c{ci}_o": o.numpy(), f"c{ci}_L": L.numpy()})
        ci += 1
    flat["n"] = np.int64(ci)
    np.savez_compressed(os.path.join(out, "triton_tree_attn.npz"), **flat)
    print("triton_tree_attn:", ci, "cases")


# ---------------------------------------------------------------------------
def _record_rounds(obj, cls):
    """Wrap tree_verification to record each round's inputs/outputs."""
    rounds = []
    raw = cls.tree_verification
    raw = getattr(raw, "__wrapped__", raw)

    def wrapped(self, input_ids, output_ids, tree_mask, cache_lens, non_leaf_len):
        res = raw(self, input_ids, output_ids, tree_mask, cache_lens, non_leaf_len)
        rounds.append(dict(all_spec=input_ids[0].numpy().copy(), llm_pred=output_ids[0].numpy().copy(),
                           tree_mask=tree_mask[0].numpy().copy(), cache_len=int(cache_lens[0]),
                           acc_ids=res[0][0].numpy().copy(), acc_num=int(res[1][0])))
        return res
    obj.tree_verification = types.MethodType(wrapped, obj)
    return rounds


def gen_e2e(out):
    """Full reference tree_spec_generate / vanilla_generate on tiny synthetic models."""
    runs = [
        ("e2e_llama_tree44", "tiny", False, [4, 4], 40, 32, 0),
        ("e2e_llama_default_tree", "tiny-gqa", False, [4, 16, 16, 16, 16], 600, 48, 1),
        ("e2e_llama_chain", "tiny", False, [1, 1, 1, 1, 1], 33, 40, 2),
        ("e2e_llama_noisy_draft", "tiny", False, [4, 8, 8], 70, 40, 3),
        ("e2e_qwen2_tree", "tiny-qwen", True, [4, 8, 8], 50, 40, 4),
    ]
    for name, preset, qwen, shape, P, gen_len, seed in runs:
        dims = PRESETS[preset]
        draft_scale = 3.0 if "noisy" in name else None
        W = synthetic_state_dict(dims, seed=seed, draft_scale=draft_scale)
        cfg = ref_shim.make_config(hidden=dims.hidden, heads=dims.heads, kv_heads=dims.kv_heads,
                                   layers=dims.layers, ffn=dims.ffn, vocab=dims.vocab,
                                   rope_theta=dims.rope_theta, qwen=qwen)
        cfg.eos_token_id = dims.vocab + 7          # never produced: fixed round count
        if qwen:
            obj = ref_shim.build_qwen2_glide(cfg, W)
            import qwen2_glide as mod
            cls = mod.Qwen2Glide
        else:
            obj = ref_shim.build_llama_glide(cfg, W)
            import llama_glide as mod
            cls = mod.LlamaGlide
        ids = synthetic_prompt(dims.vocab, P, seed=seed)
        plen = torch.tensor([P])
        eos_arg = dims.vocab + 9
        with torch.inference_mode():
            v_out, v_num, _ = obj.vanilla_generate(ids, plen, max_gen_len=gen_len, eos_id=eos_arg)
            rounds = _record_rounds(obj, cls)
            t_out, count, num, _, _ = obj.tree_spec_generate(ids, plen, tree_shape=shape, max_gen_len=gen_len,
                                                             eos_id=eos_arg, temperature=0.0)
        n_tok = int(count) + int(num)
        assert torch.equal(t_out[0, :n_tok], v_out[0, :n_tok]), "reference tree != vanilla"
        flat = dict(preset=np.array(preset), qwen=np.int64(qwen), tree_shape=np.array(shape), seed=np.int64(seed),
                    draft_scale=np.float64(draft_scale or 0.0), prompt=ids[0].numpy(), max_gen_len=np.int64(gen_len),
                    eos_arg=np.int64(eos_arg), config_eos=np.int64(cfg.eos_token_id),
                    vanilla_out=v_out[0].numpy(), vanilla_num=np.int64(v_num),
                    tree_out=t_out[0].numpy(), count=np.int64(int(count)), num=np.int64(num),
                    n_rounds=np.int64(len(rounds)),
                    w_checksum=np.float64(sum(float(v.double().abs().sum()) for v in W.values())))
        for i, r in enumerate(rounds):
            for k, v in r.items():
                flat[f"r{i}_{k}"] = np.asarray(v)
        np.savez_compressed(os.path.join(out, name + ".npz"), **flat)
        print(f"{name}: tokens={n_tok} count={int(count)} num={num} tau={(n_tok) / num:.2f} rounds={len(rounds)}")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--triton", action="store_true")
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    torch.set_num_threads(8)
    if args.triton:
        gen_triton(HERE)
    else:
        for fn in (gen_tree_verification, gen_target_tree_part, gen_draft_tree_part, gen_e2e):
            if args.only and args.only not in fn.__name__:
                continue
            fn(HERE)
