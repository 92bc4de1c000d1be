"""Generate tests/golden/ref_vectors.npz by RUNNING THE REFERENCE's own Python implementation.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python oracle/make_golden.py

Producers (the functions SURVEY.md §8c found trustworthy; everything else in the reference is known-defective):
  * fa1_forward_torch / fa1_backward_torch   reference src/fa1/torch/impl.py:26-115   (causal + non-causal)
  * fa3_forward_torch                        reference src/fa3/torch/impl.py:74-116   (fp8=False)
  * reference_attention / reference_backward reference src/common/correctness.py:5-34 (NON-causal only: D1)
Inputs follow the reference's tests: torch.manual_seed(seed) then q, k, v = randn (tests/utils.py:7-16), do =
randn_like(o); shapes/seeds from tests/test_correctness_fa{1,2,3}.py, plus multi-tile and head-dim-128 cases.
Inputs are stored next to the outputs so the fixtures do not depend on torch's RNG stream.
"""
from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np
import torch

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parents[1] / "tests" / "golden"

# (name, seed, (B,H,N,D), dtype, tag of the reference test the shape/seed comes from)
CASES = [
    ("fa1_fwd_a", 0, (1, 2, 16, 32), "float16", "tests/test_correctness_fa1.py:12-33"),
    ("fa1_fwd_b", 0, (2, 1, 33, 64), "float32", "tests/test_correctness_fa1.py:12-33"),
    ("fa1_bwd", 1, (1, 2, 12, 32), "float32", "tests/test_correctness_fa1.py:36-53"),
    ("fa1_cuda", 3, (1, 2, 24, 64), "float16", "tests/test_correctness_fa1.py:84-110"),
    ("fa2_fwd_a", 10, (1, 1, 24, 32), "float16", "tests/test_correctness_fa2.py:12-33"),
    ("fa2_fwd_b", 10, (2, 2, 33, 64), "float32", "tests/test_correctness_fa2.py:12-33"),
    ("fa2_bwd", 11, (1, 2, 16, 40), "float32", "tests/test_correctness_fa2.py:36-53"),
    ("fa2_cuda", 13, (1, 2, 32, 48), "float16", "tests/test_correctness_fa2.py:84-110"),
    ("fa3_fwd", 20, (1, 2, 24, 32), "float16", "tests/test_correctness_fa3.py:12-34"),
    ("fa3_cuda", 22, (1, 2, 32, 32), "float16", "tests/test_correctness_fa3.py:65-92"),
    ("multi_tile_d64", 100, (1, 1, 300, 64), "float16", "extra: 3 row tiles x 3 col tiles, ragged"),
    ("multi_tile_d128", 101, (1, 1, 257, 128), "bfloat16", "extra: head dim 128, ragged"),
]


def block_sparse_golden():
    """tests/golden/ref_block_sparse.npz: outputs of the reference's stand-alone attention module
    (src/fa3/torch/flashattention_pytorch.py) in eval mode -- its block-sparse branch (`_block_sparse_flash_attention`,
    :94-174) and its dense branch (:80-87) -- on the same q, k, v, so the oracle's masked attention
    (dense_ext_backward_fp32 with dropout 0) is pinned to the reference's own numbers.  The module file imports
    datasets / tiktoken at the top (not installed here), so only its `MultiHeadAttention` class is executed: the class
    source is read from the reference tree and exec'd with torch / nn / math in scope (nothing is copied)."""
    import math

    from torch import nn

    src = (REF / "src" / "fa3" / "torch" / "flashattention_pytorch.py").read_text()
    start = src.index("class MultiHeadAttention(nn.Module):")
    end = src.index("def look_ahead_mask_")
    scope = {"torch": torch, "nn": nn, "math": math}
    exec(compile(src[start:end], "flashattention_pytorch.py::MultiHeadAttention", "exec"), scope)
    mha_cls = scope["MultiHeadAttention"]
    arrays, cases = {}, []
    for name, seed, (b, h, n, d), block in (("bs_a", 200, (1, 2, 64, 16), 16), ("bs_b", 201, (2, 2, 80, 32), 16),
                                            ("bs_c", 202, (1, 1, 96, 64), 32)):
        torch.manual_seed(seed)
        q, k, v = (torch.randn(b, h, n, d) for _ in range(3))
        nb = (n + block - 1) // block
        bm = (torch.rand(nb, nb) < 0.5).to(torch.int32)
        bm[torch.arange(nb), torch.arange(nb)] = 1  # every query block keeps its diagonal block
        mod = mha_cls(d_model=h * d, num_heads=h, dropout=0.0, block_size=block).eval()
        with torch.no_grad():
            o_sparse = mod._block_sparse_flash_attention(q, k, v, 1.0, None, bm)
            causal = torch.tril(torch.ones(n, n))[None, None]
            o_sparse_causal = mod._block_sparse_flash_attention(q, k, v, 1.0, causal, bm)
            # the dense branch, written out as the module does (:80-87) with the element mask the tiles imply
            elem = bm.repeat_interleave(block, 0).repeat_interleave(block, 1)[:n, :n][None, None]
            scores = torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(d)
            o_dense = torch.matmul(torch.softmax(scores.masked_fill(elem == 0, float("-inf")), dim=-1), v)
        for key, t in (("q", q), ("k", k), ("v", v), ("block_mask", bm), ("o_sparse", o_sparse),
                       ("o_sparse_causal", o_sparse_causal), ("o_dense_masked", o_dense)):
            arrays[f"{name}/{key}"] = t.numpy()
        cases.append({"name": name, "seed": seed, "shape": [b, h, n, d], "block": block})
    np.savez_compressed(OUT / "ref_block_sparse.npz", **arrays)
    (OUT / "ref_block_sparse.json").write_text(json.dumps({
        "generator": "oracle/make_golden.py::block_sparse_golden", "torch": torch.__version__,
        "reference_functions": ["fa3.torch.flashattention_pytorch.MultiHeadAttention._block_sparse_flash_attention (eval)"],
        "cases": cases}, indent=1))
    print(f"wrote ref_block_sparse.npz: {len(arrays)} arrays, {(OUT / 'ref_block_sparse.npz').stat().st_size / 1024:.0f} KiB")


def fp8_helpers_golden():
    """tests/golden/ref_fp8_helpers.npz: the one piece of the reference's fp8 emulation that is sound on its own — the
    per-block absolute maxima of `_block_absmax_scale` (src/fa3/torch/impl.py:20-31), ragged last block included.  The
    oracle's per-block scales are these maxima / 448.  (The emulation's Hadamard step is not a Walsh-Hadamard transform
    and its quantiser only clamps to [-1, 1] — SURVEY.md D5 — so nothing else of it can serve as a pin.)"""
    sys.path.insert(0, str(REF / "src"))
    from fa3.torch.impl import _block_absmax_scale  # noqa: E402

    arrays, cases = {}, []
    for name, seed, (bh, n, d), block in (("amax_a", 300, (2, 300, 128), 128), ("amax_b", 301, (3, 128, 128), 128),
                                          ("amax_c", 302, (1, 77, 128), 128)):
        torch.manual_seed(seed)
        x = torch.randn(bh, n, d) * torch.rand(bh, n, 1) * 4.0
        arrays[f"{name}/x"] = x.numpy()
        arrays[f"{name}/block_absmax"] = _block_absmax_scale(x, block).numpy()
        cases.append({"name": name, "seed": seed, "shape": [bh, n, d], "block": block})
    np.savez_compressed(OUT / "ref_fp8_helpers.npz", **arrays)
    (OUT / "ref_fp8_helpers.json").write_text(json.dumps({
        "generator": "oracle/make_golden.py::fp8_helpers_golden", "torch": torch.__version__,
        "reference_functions": ["fa3.torch.impl._block_absmax_scale"], "cases": cases}, indent=1))
    print(f"wrote ref_fp8_helpers.npz: {len(arrays)} arrays")


def main():
    sys.path.insert(0, str(REF / "src"))
    from common.correctness import reference_attention, reference_backward  # noqa: E402
    from fa1.spec import pick_fa1_spec  # noqa: E402
    from fa1.torch.impl import fa1_backward_torch, fa1_forward_torch  # noqa: E402
    from fa3.torch.impl import fa3_forward_torch  # noqa: E402

    arrays, manifest = {}, []

    def put(key, t):
        t = t.detach()
        arrays[key] = (t.float() if t.dtype == torch.bfloat16 else t).numpy()  # npz has no bf16: store exactly as fp32

    for name, seed, (b, h, n, d), dtype_name, origin in CASES:
        dtype = getattr(torch, dtype_name)
        torch.manual_seed(seed)
        q, k, v = (torch.randn((b, h, n, d), dtype=dtype).reshape(b * h, n, d) for _ in range(3))
        do = torch.randn((b * h, n, d), dtype=dtype)
        scale = d ** -0.5
        spec = pick_fa1_spec(d)
        for t, key in ((q, "q"), (k, "k"), (v, "v"), (do, "do")):
            put(f"{name}/{key}", t)
        for causal in (False, True):
            tag = f"{name}/{'causal' if causal else 'full'}"
            o, lse = fa1_forward_torch(q, k, v, causal, scale, spec.br, spec.bc)
            dq, dk, dv = fa1_backward_torch(q, k, v, o, do, lse, causal, scale, spec.br, spec.bc)
            for t, key in ((o, "o"), (lse, "lse"), (dq, "dq"), (dk, "dk"), (dv, "dv")):
                put(f"{tag}/{key}", t)
            if name.startswith("fa3"):
                o3, lse3 = fa3_forward_torch(q, k, v, causal, scale, spec.br, spec.bc)
                put(f"{tag}/o_fa3", o3)
                put(f"{tag}/lse_fa3", lse3)
        # dense reference, non-causal only (the causal branch of reference_attention is broken: SURVEY.md D1)
        o_d, lse_d = reference_attention(q, k, v, causal=False, softmax_scale=scale)
        dq_d, dk_d, dv_d, _, _ = reference_backward(q, k, v, do, False, scale)
        for t, key in ((o_d, "o"), (lse_d, "lse"), (dq_d, "dq"), (dk_d, "dk"), (dv_d, "dv")):
            put(f"{name}/dense_full/{key}", t)
        manifest.append({"name": name, "seed": seed, "bh": b * h, "n": n, "d": d, "dtype": dtype_name,
                         "scale": scale, "br": spec.br, "bc": spec.bc, "origin": origin})

    OUT.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(OUT / "ref_vectors.npz", **arrays)
    (OUT / "ref_vectors.json").write_text(json.dumps({
        "generator": "oracle/make_golden.py", "torch": torch.__version__,
        "reference_functions": ["fa1.torch.impl.fa1_forward_torch", "fa1.torch.impl.fa1_backward_torch",
                                "fa3.torch.impl.fa3_forward_torch", "common.correctness.reference_attention(causal=False)",
                                "common.correctness.reference_backward(causal=False)"],
        "cases": manifest}, indent=1))
    size = (OUT / "ref_vectors.npz").stat().st_size
    print(f"wrote {len(arrays)} arrays, {size / 1024:.0f} KiB")
    block_sparse_golden()
    fp8_helpers_golden()


if __name__ == "__main__":
    main()
