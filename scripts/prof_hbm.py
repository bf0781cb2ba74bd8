"""Micro-driver for profiling the HBM-bound kernels at the sizes the training step runs them (upscale8, B = 8):
nearest upsample fwd/bwd (128 ch, 64x64x10 -> 128x128x10), the fused wind-loss forward / backward, the x-unfold of
hr_convs.2's gradient, the multi-tensor Adam step over the generator's 35.2 M parameters, instance noise, the fused
validation metrics and the input pipeline.  Prints CUDA-event times and achieved GB/s against the algorithmic bytes.
Usage: python scripts/prof_hbm.py [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes as C
import torch
from gan_sr_wind_field_b200 import _lib, ops
from gan_sr_wind_field_b200.optim import WindAdam

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
dev = "cuda"
ops.set_precision("bf16")
B, X, Z = 8, 128, 10
g = torch.Generator(device=dev).manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2


def timed(name, fn, nbytes):
    fn()
    torch.cuda.synchronize()
    ms = 0.0
    for _ in range(reps):
        flush.zero_()  # evict the operands from L2 between repetitions
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ms += e0.elapsed_time(e1)
    ms /= reps
    print(f"{name:34s} {ms * 1e3:8.1f} us  {nbytes / 1e6:8.1f} MB algorithmic  {nbytes / ms / 1e6:8.1f} GB/s", flush=True)


# nearest upsample (bf16 channels-last, the last UpConv stage) and its backward
a = ops.empty_cl(B, 128, X // 2, X // 2, Z, torch.bfloat16, dev); a.copy_(torch.randn(a.shape, device=dev, generator=g))
up = ops.empty_cl(B, 128, X, X, Z, torch.bfloat16, dev)
timed("upsample_fwd (64^2 -> 128^2)", lambda: ops.upsample_fwd(a, up), a.numel() * 2 * 5)
din = ops.empty_cl(B, 128, X // 2, X // 2, Z, torch.bfloat16, dev)
timed("upsample_bwd (128^2 -> 64^2)", lambda: ops.upsample_bwd(up, din), a.numel() * 2 * 5)

# fused wind loss
HR = torch.rand((B, 3, X, X, Z), device=dev, generator=g) * 2 - 1
SR = (HR + 0.05 * torch.randn(HR.shape, device=dev, generator=g)).requires_grad_(True)
Zt = torch.rand((B, 1, X, X, 1), device=dev, generator=g) * 400 + torch.linspace(2, 68, Z, device=dev).reshape(1, 1, 1, 1, Z)
Zt = Zt.contiguous()
xs = torch.cumsum(torch.full((X,), 200.0, device=dev), 0)
vox = B * X * X * Z
timed("windloss fwd (fused)", lambda: ops.windloss_slots(HR, SR.detach(), Zt, xs, xs), vox * 28)


def wl_bwd():
    s = ops.windloss_slots(HR, SR, Zt, xs, xs)
    (s[0] + s[2] + s[4]).backward()
    SR.grad = None


timed("windloss fwd + bwd", wl_bwd, vox * (28 + 28 + 12))

# x-unfold of the hr_convs.2 output gradient (3 channels fp32 NCXYZ -> 16 channels bf16 channels-last)
dout = torch.randn((B, 3, X, X, Z), device=dev, generator=g)
u = ops.empty_cl(B, 16, X, X, Z, torch.bfloat16, dev)


def xunfold():
    dv, uv = _lib.view(dout), _lib.view(u)
    _lib.check(_lib.load().ws_xunfold(C.byref(dv), C.byref(uv), B, 3, 5, 2, 16, X, X, Z, _lib.stream_ptr()), "xunfold")


timed("xunfold (3 -> 16 ch)", xunfold, vox * (3 * 4 + 16 * 2))

# Adam over the generator's parameter shapes
from gan_sr_wind_field_b200.CNN_models.Generator_3D_Resnet_ESRGAN import Generator_3D
G = Generator_3D(4, 3, 128, 16, upscale=8, hr_kern_size=5, lff_kern_size=1, dropout_probability=0.1).to(dev)
opt = WindAdam(G.parameters(), lr=8e-5)
for p in G.parameters():
    p.grad = torch.randn_like(p)
nparam = sum(p.numel() for p in G.parameters())
timed("adam_multi (35.2 M params)", lambda: opt.step(), nparam * 4 * 7)

# instance noise, validation metrics, input pipeline
timed("instance noise (HR batch)", lambda: ops.add_instance_noise(HR, 1.0), HR.numel() * 8)
LR = torch.rand((B, 4, X // 8, X // 8, Z), device=dev, generator=g)
timed("validation metrics (fused)", lambda: ops.validation_metrics(HR, SR.detach(), LR), HR.numel() * 8)
f64 = [torch.randn((B, X, X, Z), device=dev, generator=g, dtype=torch.float64) for _ in range(4)]
aug = torch.tensor([[0, 0, i % 4, i % 2, (i // 2) % 2] for i in range(B)], dtype=torch.int32, device=dev)
timed("prepare_batch (crop+norm+aug)", lambda: ops.prepare_batch(f64[0], f64[1], f64[2], f64[3], aug=aug, coarseness=8,
                                                                  include_z_channel=True, uvw_max=32.33, z_min=-2.71,
                                                                  z_max=550.44),
      vox * (4 * 8 + 4 * 4) + B * 4 * (X // 8) ** 2 * Z * 4)
print("ok")
