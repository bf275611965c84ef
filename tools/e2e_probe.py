import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "stable-diffusion-xl-burn_b200")):
    sys.path.insert(0, p)
import torch, sdxl_b200
sys.path.insert(0, ROOT)
import bench
ctx = sdxl_b200.Context(0)
cfg = sdxl_b200.SDXL_BASE
d = sdxl_b200.Diffuser(ctx, cfg, sdxl_b200.build_pack(sdxl_b200.synth_weights(cfg, seed=0, device="cuda:0")))
cond = bench.make_conditioning(0, torch.device("cuda", 0))
d.sampler_begin(cond, 7.5)
d.sampler_set_latent(torch.randn(1, 4, 128, 128))
for _ in range(4):
    d.sampler_step(999, 966)
ctx.synchronize()
def t(f, n=8):
    ctx.synchronize(); w0 = time.perf_counter()
    for _ in range(n): f()
    ctx.synchronize(); return (time.perf_counter() - w0) * 1e3 / n
print("async steps ms", t(lambda: d.sampler_step(999, 966)))
print("step + sync ms", t(lambda: (d.sampler_step(999, 966), ctx.synchronize())))
host = torch.randn(1, 4, 128, 128).pin_memory()
print("step_host ms", t(lambda: d.sampler_step_host(999, 966, host)))
host2 = torch.randn(1, 4, 128, 128)
print("step_host pageable ms", t(lambda: d.sampler_step_host(999, 966, host2)))
