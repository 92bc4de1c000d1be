"""GPU: the real FP8 (e4m3) forward behind fa3's fp8=True against this repo's own
quantise -> dequantise fp32 oracle (the reference's fp8 emulation is broken and pins nothing).
Tolerances: the quantisation pre-pass is checked bit for bit (e4m3 bytes) and to fp32 rounding (scales).  The forward:
LSE depends only on the (exactly accumulated) e4m3 products, so it is held to 2e-3 against fp32 maths on the quantised
inputs; O additionally carries the e4m3 rounding of the probabilities (3 mantissa bits: relative 2^-4 per element, not
modelled by the oracle because it depends on the running row maximum), so it is held to an RMS error of 1e-2 and a
max-abs error of 2^-4 * max|V| (+ 2e-2); measured values are printed."""
import pytest
import torch

import flashattention_lab_cuda as ext
import probes
from fa3 import fa3_attention
from oracle.attention_oracle import dense_forward, error_report, fp8_forward_oracle, fp8_quantize_dequantize

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", [6, 7])
def test_e4m3_umma_descriptors(mode):
    torch.manual_seed(mode)
    a = torch.randn(128, 128, device="cuda").to(torch.float8_e4m3fn)
    b = torch.randn(128, 128, device="cuda").to(torch.float8_e4m3fn)
    out = probes.probe_umma(mode, a, b)
    want = a.float() @ (b.float().T if mode == 6 else b.float())
    assert (out - want).abs().max() < 1e-3 * want.abs().max()  # exact products, fp32 accumulation order only


@pytest.mark.parametrize("hadamard", [False, True])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_quantisation_prepass_matches_the_oracle(hadamard, dtype):
    torch.manual_seed(1)
    x = torch.randn(3, 300, 128, device="cuda", dtype=dtype) * torch.linspace(0.1, 4.0, 300, device="cuda")[None, :, None].to(dtype)
    q8, scales = ext.fp8_quantize_raw(x, hadamard)
    deq, q8_ref, scale_ref = fp8_quantize_dequantize(x, hadamard)
    assert torch.allclose(scales.cpu(), scale_ref, rtol=1e-5, atol=0)
    got = q8.cpu().view(torch.float8_e4m3fn).float()
    want = q8_ref.float()
    # identical up to fp32 summation order in the Hadamard butterflies: at most a handful of values may land on the
    # neighbouring e4m3 code
    mismatch = (got != want).float().mean().item()
    assert mismatch < (2e-3 if hadamard else 1e-6), mismatch
    assert ((got - want).abs() <= 0.13 * want.abs().clamp_min(2 ** -6)).all()


@pytest.mark.parametrize("bh,n,causal,dtype", [(2, 256, False, torch.bfloat16), (3, 1000, True, torch.bfloat16),
                                               (2, 2048, True, torch.float16), (1, 4096, False, torch.bfloat16),
                                               (2, 333, True, torch.bfloat16)])
def test_fp8_forward(bh, n, causal, dtype):
    torch.manual_seed(n)
    d = 128
    q, k, v = (torch.randn(bh, n, d, device="cuda", dtype=dtype) for _ in range(3))
    scale = d ** -0.5
    o, lse = ext.fa3_forward(q, k, v, causal, scale, 64, 128, 2, True)
    assert o.dtype == dtype and o.shape == q.shape and lse.dtype == torch.float32
    o_q, lse_q = fp8_forward_oracle(q, k, v, causal, scale)          # fp32 maths on the quantised inputs
    o_x, lse_x = dense_forward(q.float().cpu(), k.float().cpu(), v.float().cpu(), causal, scale)  # exact inputs
    rep_l = error_report(lse, lse_q, 2e-3, 2e-3)
    diff = o.float().cpu() - o_q.float()
    rms, mx = diff.pow(2).mean().sqrt().item(), diff.abs().max().item()
    bound = 2 ** -4 * v.float().abs().max().item() + 2e-2
    print(f"fp8 n={n} causal={causal}: o - oracle_q rms {rms:.3e} max {mx:.3e} (bound {bound:.3e}), |lse - oracle_q| max "
          f"{rep_l['max_abs']:.3e}; quantisation itself moves o by max {(o_q - o_x).abs().max().item():.3e} "
          f"rms {(o_q - o_x).float().pow(2).mean().sqrt().item():.3e}")
    assert rep_l["violations"] == 0, rep_l
    assert rms < 1e-2 and mx < bound
    assert torch.isfinite(o.float()).all()


def test_fp8_through_the_public_entry_point_with_backward():
    torch.manual_seed(5)
    b, h, n, d = 1, 2, 512, 128
    q, k, v = (torch.randn(b, h, n, d, device="cuda", dtype=torch.bfloat16).requires_grad_(True) for _ in range(3))
    o, lse = fa3_attention(q, k, v, causal=True, backend="cuda", fp8=True)
    o.backward(torch.randn_like(o))
    o16, _ = fa3_attention(q.detach(), k.detach(), v.detach(), causal=True, backend="cuda")
    assert (o.float() - o16.float()).pow(2).mean().sqrt() < 5e-2  # the fp8 forward tracks the 16-bit one
    for t in (q, k, v):
        assert t.grad is not None and torch.isfinite(t.grad.float()).all()
    with pytest.raises(NotImplementedError):
        x = torch.randn(1, 2, 64, 64, device="cuda", dtype=torch.bfloat16)
        fa3_attention(x, x, x, backend="cuda", fp8=True)
