import sys
sys.path[:0] = ["flashattention-pytorch_b200", "."]
import torch, probes
torch.manual_seed(0)
a = (torch.randn(128, 128, device="cuda")).to(torch.float8_e4m3fn)
b = (torch.randn(128, 128, device="cuda")).to(torch.float8_e4m3fn)
af, bf = a.float(), b.float()
for mode, want in ((6, af @ bf.T), (7, af @ bf)):
    out = probes.probe_umma(mode, a, b)
    torch.cuda.synchronize()
    print(f"fp8 probe mode {mode}: max_abs_err={(out - want).abs().max().item():.4e} ref_max={want.abs().max().item():.2f}", flush=True)
