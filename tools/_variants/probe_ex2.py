import sys
sys.path[:0] = ["flashattention-pytorch_b200", "."]
import torch, probes
sms = torch.cuda.get_device_properties(0).multi_processor_count
for mode, name, per in ((0, "ex2.f32", 1), (1, "ex2.f16x2", 2), (2, "ex2.bf16x2", 2)):
    iters, ctas = 20000, sms * 8
    probes.probe_ex2_rate(mode, 100, ctas); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); probes.probe_ex2_rate(mode, iters, ctas); b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    instr = ctas * 256 * iters * 8
    print(f"{name}: {ms:.3f} ms, {instr / ms / 1e6 / sms:.1f} Ginstr/s/SM = {instr / (ms * 1e-3) / sms / 1.965e9:.2f} thread-instr/clk/SM at 1965 MHz, "
          f"{instr * per / (ms * 1e-3) / sms / 1.965e9:.2f} results/clk/SM", flush=True)
