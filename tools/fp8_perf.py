"""FP8 (e4m3) forward rate next to the bf16 forward at the same shapes (d = 128): the e4m3 kernel alone on pre-quantised
inputs, the quantisation pre-pass, and the whole fa3 fp8=True call.  usage (under gpurun): python tools/fp8_perf.py"""
import ctypes
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT / "flashattention-pytorch_b200"), str(ROOT)]
import torch
import flashattention_lab_cuda as ext


def t_ms(fn, it=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / it


lib = ext.load_library()
for bh, n, causal in ((64, 4096, True), (64, 8192, False), (64, 8192, True), (16, 16384, True)):
    d = 128
    q, k, v = (torch.randn(bh, n, d, device="cuda", dtype=torch.bfloat16) for _ in range(3))
    scale = d ** -0.5
    f = 4.0 * bh * n * n * d * (0.5 if causal else 1.0)
    q8, sq = ext.fp8_quantize_raw(q, True)
    k8, sk = ext.fp8_quantize_raw(k, True)
    v8, sv = ext.fp8_quantize_raw(v, False)
    sv_ref = sv.amax(dim=1).contiguous()
    o = torch.empty_like(q)
    lse = torch.empty(bh, n, device="cuda", dtype=torch.float32)
    shape = ext.make_shape(bh, n, n, d, 1, causal, scale)
    stream = torch.cuda.current_stream().cuda_stream

    def kernel_only():
        rc = lib.fa_sm100_fwd_fp8(ctypes.byref(shape), q8.data_ptr(), k8.data_ptr(), v8.data_ptr(), sq.data_ptr(), sk.data_ptr(),
                                  sv.data_ptr(), sv_ref.data_ptr(), o.data_ptr(), lse.data_ptr(), stream)
        assert rc == 0, rc

    t_k = t_ms(kernel_only)
    t_q = t_ms(lambda: (ext.fp8_quantize_raw(q, True), ext.fp8_quantize_raw(k, True), ext.fp8_quantize_raw(v, False)))
    t_all = t_ms(lambda: ext.fwd_fp8_raw(q, k, v, causal, scale))
    t_16 = t_ms(lambda: ext.fwd_raw(q, k, v, causal, scale))
    print(f"bh={bh} n={n} causal={int(causal)}: e4m3 kernel {t_k:.3f} ms {f / t_k / 1e9:7.1f} TFLOP/s | quantise q,k,v {t_q:.3f} ms | "
          f"fp8 call {t_all:.3f} ms {f / t_all / 1e9:7.1f} | bf16 forward {t_16:.3f} ms {f / t_16 / 1e9:7.1f} TFLOP/s", flush=True)
