"""Block-sparse and dropout variants next to the dense kernels (bf16, d = 128, N = 8192): a banded (sliding-window) tile
mask at several widths -- time should follow the number of active tiles -- and dropout p = 0.1.
usage (under gpurun): python tools/ext_perf.py"""
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


bh, n, d = 64, 8192, 128
scale = d ** -0.5
q, k, v, do = (torch.randn(bh, n, d, device="cuda", dtype=torch.bfloat16) for _ in range(4))
nb = n // 128
o, lse = ext.fwd_raw(q, k, v, True, scale)
t_f = t_ms(lambda: ext.fwd_raw(q, k, v, True, scale))
t_b = t_ms(lambda: ext.bwd_raw(q, k, v, o, do, lse, True, scale))
f_dense = 4.0 * bh * n * n * d * 0.5
print(f"dense causal: fwd {t_f:.3f} ms ({f_dense / t_f / 1e9:.0f} TFLOP/s), bwd {t_b:.3f} ms ({2.5 * f_dense / t_b / 1e9:.0f})", flush=True)
idx = torch.arange(nb)
for width in (nb, 16, 8, 4):
    mask = ((idx[:, None] - idx[None, :]).abs() < width).to(torch.uint8).cuda()  # banded: |i - j| < width tiles
    active = int(torch.tril(mask.cpu()).sum())  # causal on top
    frac = active / (nb * (nb + 1) / 2)
    o_s, lse_s = ext.fwd_ex_raw(q, k, v, True, scale, block_mask=mask)
    tf = t_ms(lambda: ext.fwd_ex_raw(q, k, v, True, scale, block_mask=mask))
    tb = t_ms(lambda: ext.bwd_ex_raw(q, k, v, o_s, do, lse_s, True, scale, block_mask=mask))
    fl = 4.0 * bh * active * 128 * 128 * d  # tiles actually computed (diagonal tiles counted in full)
    print(f"band {width:3d} tiles ({100 * frac:5.1f}% of the causal tiles): fwd {tf:.3f} ms ({t_f / tf:4.1f}x dense, {fl / tf / 1e9:.0f} TFLOP/s "
          f"on computed tiles), bwd {tb:.3f} ms ({t_b / tb:4.1f}x dense, {2.5 * fl / tb / 1e9:.0f})", flush=True)
o_d, lse_d = ext.fwd_ex_raw(q, k, v, True, scale, dropout_p=0.1, seed=1)
tf = t_ms(lambda: ext.fwd_ex_raw(q, k, v, True, scale, dropout_p=0.1, seed=1))
tb = t_ms(lambda: ext.bwd_ex_raw(q, k, v, o_d, do, lse_d, True, scale, dropout_p=0.1, seed=1))
print(f"dropout 0.1: fwd {tf:.3f} ms ({f_dense / tf / 1e9:.0f} TFLOP/s, {tf / t_f:.2f}x the dense time), bwd {tb:.3f} ms "
      f"({2.5 * f_dense / tb / 1e9:.0f}, {tb / t_b:.2f}x)", flush=True)
