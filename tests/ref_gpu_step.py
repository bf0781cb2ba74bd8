"""Context number (NOT the product path, NOT bench.py's reference arm): the reference's G step restated with stock
torch ops (oracle/wind_oracle.py) executed on the B200 through cuDNN, in fp32 (TF32 allowed, torch's default for
convs) and under bf16 autocast — i.e. what a user of the reference gets on this GPU today.
Usage: python tests/ref_gpu_step.py [batch]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from gan_sr_wind_field_b200.CNN_models.Discriminator_3D import Discriminator_3D
from gan_sr_wind_field_b200.CNN_models.Generator_3D_Resnet_ESRGAN import Generator_3D
from gan_sr_wind_field_b200.config.config import Config
from gan_sr_wind_field_b200.tools import initialization
from oracle import wind_oracle as wo

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda:0")
cfg = Config(bench.INI)
g, d, t = cfg.generator, cfg.discriminator, cfg.training
torch.manual_seed(2001)
G = Generator_3D(4, 3, g.num_features, g.num_RRDB, upscale=cfg.scale, hr_kern_size=g.hr_kern_size,
                 lff_kern_size=g.lff_kern_size, dropout_probability=g.dropout_probability)
initialization.init_weights(G, g.weight_init_scale)
D = Discriminator_3D(3, d.num_features)
initialization.init_weights(D, d.weight_init_scale)
pG = {k: v.detach().to(dev).requires_grad_(v.is_floating_point()) for k, v in G.state_dict().items()}
pD = {k: v.detach().to(dev) for k, v in D.state_dict().items()}
opt = torch.optim.Adam([v for v in pG.values() if v.requires_grad], lr=t.learning_rate_g)
LR, HR, Z, x, y = (v.to(dev) for v in wo.synthetic_batch(B, 128, 10, 8, seed=2001))
w = dict(pixel=t.pixel_loss_weight, xy=t.gradient_xy_loss_weight, z=t.gradient_z_loss_weight,
         div=t.divergence_loss_weight, dxy=t.xy_divergence_loss_weight, adv=t.adversarial_loss_weight)
real, fake = torch.full((B,), 0.9, device=dev), torch.zeros(B, device=dev)
torch.backends.cudnn.benchmark = True

def step(autocast):
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        SR = wo.generator_forward(pG, LR, Z)
        with torch.no_grad():
            y_pred = wo.discriminator_forward(pD, HR, False).reshape(-1)
        y_fake = wo.discriminator_forward(pD, SR, False).reshape(-1)
    SR = SR.float()
    adv = wo.adversarial_G(y_pred.float(), y_fake.float(), real, fake, t.gan_type)
    total, _ = wo.generator_loss(HR, SR, Z, x, y, w, adv=adv)
    total.backward()
    opt.step()

for name, ac, tf32 in (("fp32 (cudnn TF32 allowed, torch default)", False, True), ("fp32 strict (TF32 off)", False, False),
                       ("bf16 autocast", True, True)):
    torch.backends.cudnn.allow_tf32 = tf32
    try:
        for _ in range(3):
            step(ac)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            step(ac)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"reference-on-GPU (torch {torch.__version__} + cuDNN {torch.backends.cudnn.version()}), B={B}, {name}: "
              f"{ms:.1f} ms/step, {B * 163840 / ms / 1e3:.2f} M voxels/s, peak mem {torch.cuda.max_memory_allocated() / 1e9:.1f} GB", flush=True)
    except Exception as e:  # noqa: BLE001
        print(f"{name}: failed: {type(e).__name__}: {str(e)[:200]}", flush=True)
