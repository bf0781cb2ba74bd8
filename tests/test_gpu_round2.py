"""GPU tests of the training-step machinery around the convolutions: the multi-tensor Adam kernel, instance noise,
fused validation metrics, the input pipeline (bit-exact), CUDA-graph replay of the G / D step, the D-skip at zero
adversarial weight, and (2 GPUs) data-parallel gradient averaging."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from tests.util import GOLDEN, ROOT, load_npz, rel_l2, sd_from

pytestmark = pytest.mark.gpu


def test_adam_kernel_matches_torch_adam():
    """ws_adam_step against torch.optim.Adam (CPU, fp32) over 4 steps: odd sizes, a 4-byte-aligned-only view, weight
    decay, and the found_inf skip (reference: wind_field_GAN_3D.py:151-162, 457-460)."""
    from gan_sr_wind_field_b200.optim import WindAdam
    torch.manual_seed(0)
    shapes = [(144, 144, 5, 5, 5), (3,), (32, 128, 3, 3, 3), (1,), (17, 5), (128,)]
    base = torch.randn(1 + sum(int(np.prod(s)) for s in shapes))
    cpu, gpu = [], []
    off = 1  # views at an odd element offset: not 16-byte aligned -> scalar path of the kernel
    flat_gpu = base.clone().cuda()
    for s in shapes:
        n = int(np.prod(s))
        cpu.append(base[off:off + n].clone().reshape(s).requires_grad_(True))
        gpu.append(torch.nn.Parameter(flat_gpu[off:off + n].view(s)))
        off += n
    for wd in (0.0, 0.01):
        o_cpu = torch.optim.Adam(cpu, lr=8e-5, betas=(0.9, 0.999), weight_decay=wd)
        o_gpu = WindAdam(gpu, lr=8e-5, betas=(0.9, 0.999), weight_decay=wd)
        start = [p.detach().clone() for p in cpu]
        for step in range(4):
            for a, b in zip(cpu, gpu):
                g = torch.randn(a.shape, generator=torch.Generator().manual_seed(100 * step + a.numel() % 97))
                a.grad, b.grad = g.clone(), g.cuda()
            o_gpu.found_inf = torch.tensor(1.0 if step == 2 else 0.0, device="cuda")
            if step != 2:
                o_cpu.step()
            o_gpu.step()
        for a, b, s0 in zip(cpu, gpu, start):
            assert rel_l2(b.detach().cpu() - s0, a.detach() - s0) <= 2e-5, a.shape
        st = o_gpu.state[gpu[0]]
        assert float(st["step"]) == 3.0  # the skipped step did not count
        sd = o_gpu.state_dict()
        assert set(sd["state"][0].keys()) == {"step", "exp_avg", "exp_avg_sq"}  # torch.optim.Adam's layout
    # a reference-format (non-fused, CPU `step`) state dict loads and steps
    o_ref = torch.optim.Adam([p.detach().clone().requires_grad_(True) for p in cpu], lr=1e-3)
    for p in o_ref.param_groups[0]["params"]:
        p.grad = torch.ones_like(p)
    o_ref.step()
    o_new = WindAdam(gpu, lr=1e-3)
    o_new.load_state_dict(o_ref.state_dict())
    for b in gpu:
        b.grad = torch.ones_like(b)
    o_new.step()
    assert float(o_new.state[gpu[0]]["step"]) == 2.0


def test_instance_noise_kernel():
    """x + U[0,1) * scale (trainingtricks.py:49-58): range, moments, fresh draws per call, device-side scale,
    reproducible after reseeding, identity gradient."""
    from gan_sr_wind_field_b200 import ops
    x = torch.zeros(2, 3, 64, 64, 10, device="cuda", requires_grad=True)
    ops.reseed_instance_noise(1234, x.device)
    ops._noise_state(x.device)
    ops.reseed_instance_noise(1234, x.device)
    a = ops.add_instance_noise(x, 2.0)
    b = ops.add_instance_noise(x, 2.0)
    assert float(a.min()) >= 0.0 and float(a.max()) < 2.0
    assert abs(float(a.mean()) - 1.0) < 5e-3 and abs(float(a.var()) - 4.0 / 12.0) < 5e-3
    assert not torch.equal(a, b)
    assert abs(float(((a - 1) * (b - 1)).mean())) < 5e-3  # successive calls are uncorrelated
    ops.reseed_instance_noise(1234, x.device)
    assert torch.equal(ops.add_instance_noise(x, 2.0), a)
    s = torch.tensor([0.5], device="cuda")
    c = ops.add_instance_noise(x.detach() + 3.0, 1.0, scale_dev=s)
    assert float(c.min()) >= 3.0 and float(c.max()) < 3.5
    a.sum().backward()
    assert torch.equal(x.grad, torch.ones_like(x))
    nan = ops.add_instance_noise(x.detach(), 1.0, scale_dev=torch.tensor([float("nan")], device="cuda"))
    assert bool(torch.isnan(nan).all())  # past niter+1 the reference's sqrt(<0) gives NaN, too


def test_validation_metrics_vs_golden_and_torch():
    from gan_sr_wind_field_b200 import ops
    z = load_npz("metrics.npz")
    HR, SR, LR = (torch.from_numpy(z[k]).cuda() for k in ("HR", "SR", "LR"))
    s = ops.validation_metrics(HR, SR, LR)
    vox = HR.shape[0] * HR.shape[2] * HR.shape[3] * HR.shape[4]
    psnr = lambda v: 10.0 * np.log10(4.0 / (float(v) / vox + 1e-8))
    assert abs(psnr(s[0]) - float(z["psnr"])) <= 1e-4
    assert abs(psnr(s[1]) - float(z["tri_psnr"])) <= 1e-4
    assert abs(float(s[2]) / (3 * vox) - float(z["tri_l1"])) <= 1e-6
    assert abs(float(s[1]) / (3 * vox) - float(z["tri_l2"])) <= 1e-6
    # shipped validation shape: HR 128x128x10 from LR 16x16x10 (scale 8), against torch's own interpolate on the GPU
    g = torch.Generator(device="cuda").manual_seed(1)
    HR = torch.rand((2, 3, 128, 128, 10), device="cuda", generator=g) * 2 - 1
    SR = HR + 0.1 * torch.randn(HR.shape, device="cuda", generator=g)
    LR = torch.rand((2, 4, 16, 16, 10), device="cuda", generator=g)
    s = ops.validation_metrics(HR, SR, LR)
    tri = torch.nn.functional.interpolate(LR[:, :3], scale_factor=(8, 8, 1), mode="trilinear", align_corners=True)
    ref = torch.stack((((HR - SR).double() ** 2).sum(), ((HR - tri).double() ** 2).sum(),
                       (HR - tri).double().abs().sum(), (HR - SR).double().abs().sum()))
    assert rel_l2(s, ref) <= 1e-6


def test_prepare_batch_bit_exact_vs_reference_golden():
    """ws_prepare_batch against the reference's reformat_to_torch + augmentation outputs: bit-exact."""
    from gan_sr_wind_field_b200 import ops
    z = load_npz("prepare_batch.npz")
    c = {k[6:]: float(z[k]) for k in z.files if k.startswith("const/")}
    size, cf = int(z["slice_size"]), int(z["coarseness"])
    f = {k: torch.from_numpy(z[k]).cuda()[None].contiguous() for k in ("u", "v", "w", "p", "z", "zag")}
    cases = z["cases"]
    for layout in sorted(set(int(r[0]) for r in cases)):
        idx = [i for i in range(len(cases)) if int(cases[i][0]) == layout]
        n = len(idx)
        rep = {k: v.expand(n, -1, -1, -1).contiguous() for k, v in f.items()}
        aug = torch.tensor([[int(v) for v in cases[i][4:9]] for i in idx], dtype=torch.int32, device="cuda")
        inc_p, inc_z, inc_ag = (bool(v) for v in cases[idx[0]][1:4])
        LR, HR, Z = ops.prepare_batch(rep["u"], rep["v"], rep["w"], rep["z"], pressure=rep["p"],
                                      z_above_ground=rep["zag"], aug=aug, crop=(size, size), coarseness=cf,
                                      include_pressure=inc_p, include_z_channel=inc_z,
                                      include_above_ground_channel=inc_ag, uvw_max=c["UVW_MAX"], p_min=c["P_MIN"],
                                      p_max=c["P_MAX"], z_min=c["Z_MIN"], z_max=c["Z_MAX"],
                                      z_above_ground_max=c["Z_ABOVE_GROUND_MAX"])
        for j, i in enumerate(idx):
            for name, t in (("LR", LR), ("HR", HR), ("Z", Z)):
                assert torch.equal(t[j].cpu(), torch.from_numpy(z[f"case{i}/{name}"])), (i, name, cases[i])


def _tiny_gan(graph: bool, seed=3, noise=True, lr=1e-6):
    from gan_sr_wind_field_b200 import ops
    from gan_sr_wind_field_b200.config.config import Config
    from gan_sr_wind_field_b200.GAN_models.wind_field_GAN_3D import wind_field_GAN_3D
    z = load_npz("gan_step.npz")
    cfg = Config(os.path.join(GOLDEN, "configs", "tiny_gan.ini"))
    cfg.is_train, cfg.gpu_id, cfg.device = True, 0, torch.device("cuda:0")
    cfg.generator.dropout_probability = 0.0  # dropout masks come from torch's RNG, whose offsets differ under capture
    cfg.training.use_instance_noise = noise
    # a small step size keeps the (deliberately hot, chaotic) tiny GAN on one trajectory: run-to-run differences of the
    # fp32 atomics in the BatchNorm statistics would otherwise be amplified into different losses within a few steps
    cfg.training.learning_rate_g = cfg.training.learning_rate_d = lr
    os.environ["WINDSR_CUDA_GRAPH"] = "1" if graph else "0"
    torch.manual_seed(seed)
    gan = wind_field_GAN_3D(cfg)
    gan.G.load_state_dict(sd_from(z, "G0/"))
    gan.D.load_state_dict(sd_from(z, "D0/"))
    ops.reseed_instance_noise(99, torch.device("cuda:0"))
    ops._noise_state(torch.device("cuda:0"))
    ops.reseed_instance_noise(99, torch.device("cuda:0"))
    LR, HR, Z, x, y = (torch.from_numpy(z[k]).cuda() for k in ("LR", "HR", "Z", "x", "y"))
    gan.feed_xy_niter(x, y, torch.tensor(cfg.training.niter, device="cuda"), 1, 1)  # G, D, G, D, ...
    return gan, (LR, HR, Z)


def test_graph_replay_matches_eager_steps():
    """16 iterations (8 G steps + 8 D steps, instance noise on, label smoothing ramp): eager vs captured-and-replayed.
    FP32 mode (deterministic kernels): the two runs see the same noise stream, labels and learning rate, so losses
    and weights must agree to rounding."""
    from gan_sr_wind_field_b200 import ops
    try:
        runs = []
        for graph in (False, True):
            gan, (LR, HR, Z) = _tiny_gan(graph)
            losses = []
            with ops.precision("fp32"):
                for it in range(1, 17):
                    gan.optimize_parameters(LR, HR, Z, it)
                    if gan.is_G_iteration(it):
                        losses.append(gan.get_G_train_loss_dict_ref()["total"].detach().clone())
                    else:
                        losses.append(gan.get_D_loss_dict_ref()["train_loss"].detach().clone())
            torch.cuda.synchronize()
            if graph:
                assert len([g for g in gan._graphs.values() if g]) == 2, gan._graphs
                assert all(g.replays >= 4 for g in gan._graphs.values())
            runs.append((torch.stack([l.reshape(()) for l in losses]).cpu(),
                         {k: v.detach().cpu().clone() for k, v in gan.G.state_dict().items()},
                         {k: v.detach().cpu().clone() for k, v in gan.D.state_dict().items()}))
        (l0, g0, d0), (l1, g1, d1) = runs
        assert torch.allclose(l0, l1, rtol=2e-3, atol=1e-5), (l0, l1)
        z = load_npz("gan_step.npz")
        start = sd_from(z, "G0/")
        moved = 0
        for k in g0:
            upd0, upd1 = g0[k] - start[k], g1[k] - start[k]
            if upd0.abs().max() > 0:
                moved += 1
                # Adam moves every weight by ~lr per step whatever the gradient's size: compare the direction of the
                # 8 accumulated steps (a replay that skipped or mangled updates would be uncorrelated)
                cos = float((upd0 * upd1).sum() / (upd0.norm() * upd1.norm()).clamp_min(1e-30))
                assert cos >= 0.98, (k, cos)
            assert rel_l2(g1[k], g0[k]) <= 1e-4, k
        assert moved >= len(g0) - 2
        for k in d0:
            # (biases start at 0, so after 8 Adam steps of 1e-6 they ARE the updates; a bias whose gradient is at
            # rounding-noise level — the last classifier bias under the symmetric RaGAN loss — random-walks by up to
            # 8e-6 in either run)
            if d0[k].is_floating_point():
                assert torch.allclose(d1[k], d0[k], rtol=1e-3, atol=1.2e-5), k
    finally:
        os.environ.pop("WINDSR_CUDA_GRAPH", None)


def test_skipping_D_at_zero_adversarial_weight_changes_nothing():
    """adversarial_loss_weight == 0: the G step without the discriminator (default) gives the same loss and the same
    generator gradients as the reference's order of operations, which runs D and multiplies its term by 0.0
    (wind_field_GAN_3D.py:487,426)."""
    from gan_sr_wind_field_b200 import ops
    out = []
    try:
        for skip in ("1", "0"):
            os.environ["WINDSR_SKIP_D_WHEN_ZERO"] = skip
            gan, (LR, HR, Z) = _tiny_gan(False)
            gan.cfg.training.adversarial_loss_weight = 0.0
            with ops.precision("fp32"):
                gan.optimize_parameters(LR, HR, Z, 2)  # a G iteration (period 1: even iterations)
            out.append((float(gan.get_G_train_loss_dict_ref()["total"]),
                        {k: p.grad.detach().clone() for k, p in gan.G.named_parameters()}))
        assert abs(out[0][0] - out[1][0]) <= 1e-6 * abs(out[1][0])
        for k in out[0][1]:
            assert rel_l2(out[0][1][k], out[1][1][k]) <= 1e-6, k
    finally:
        os.environ.pop("WINDSR_SKIP_D_WHEN_ZERO", None)


_DDP_GPU_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device(f"cuda:{rank}")
dist.init_process_group("nccl", device_id=dev)
from gan_sr_wind_field_b200 import ops
from gan_sr_wind_field_b200.config.config import Config
from gan_sr_wind_field_b200.GAN_models.wind_field_GAN_3D import wind_field_GAN_3D
from tests.util import GOLDEN, load_npz, rel_l2, sd_from
os.environ["WINDSR_CUDA_GRAPH"] = "0"
z = load_npz("gan_step.npz")

def make(distributed, adv=None):
    cfg = Config(os.path.join(GOLDEN, "configs", "tiny_gan.ini"))
    cfg.is_train, cfg.gpu_id, cfg.device = True, rank, dev
    cfg.generator.dropout_probability = 0.0
    cfg.training.use_instance_noise = False
    if adv is not None:
        cfg.training.adversarial_loss_weight = adv
    torch.manual_seed(100 + rank)               # DIFFERENT seeds per rank: the constructor must broadcast rank 0's weights
    gan = wind_field_GAN_3D(cfg)
    if not distributed:
        gan.sync_G = gan.sync_D = None
        gan.world_size = 1
    return gan, cfg

ops.set_precision("fp32")
LR, HR, Z, x, y = (torch.from_numpy(z[k]).to(dev) for k in ("LR", "HR", "Z", "x", "y"))
# global batch = the fixture's 2 samples repeated with a perturbation; rank r takes samples [2r, 2r+2)
g = torch.Generator().manual_seed(0)
LRg = torch.cat([LR.cpu() + 0.01 * i * torch.randn(LR.shape, generator=g) for i in range(world)]).to(dev)
HRg = torch.cat([HR.cpu() * (1.0 + 0.5 * i) + 0.01 * i * torch.randn(HR.shape, generator=g) for i in range(world)]).to(dev)
Zg = torch.cat([Z.cpu() for i in range(world)]).to(dev)
b = LR.shape[0]
sl = slice(rank * b, (rank + 1) * b)
res = {}

# ---- 1. generator step of the shipped kind (adversarial weight 0: D is not on the path).  The only batch-coupled
#         terms are the loss normalisers (global maxima), which are MAX-all-reduced: W ranks must reproduce the
#         single-process step on the CONCATENATED batch.
gan, cfg = make(True, adv=0.0)
w0 = [p.detach().clone() for p in gan.G.parameters()]
chk = torch.stack([p.double().sum() for p in w0]).sum()
chks = [torch.zeros_like(chk) for _ in range(world)]
dist.all_gather(chks, chk)
assert all(bool(c == chks[0]) for c in chks), "replicas start from different weights"
gan.feed_xy_niter(x, y, torch.tensor(100, device=dev), 1, 1)
sdG = {k: v.detach().clone() for k, v in gan.G.state_dict().items()}
sdD = {k: v.detach().clone() for k, v in gan.D.state_dict().items()}
gan.optimize_parameters(LRg[sl], HRg[sl], Zg[sl], 2)
mine = {k: p.grad.detach().clone() for k, p in gan.G.named_parameters() if p.grad is not None}
ref, _ = make(False, adv=0.0)
ref.G.load_state_dict(sdG); ref.D.load_state_dict(sdD)
ref.feed_xy_niter(x, y, torch.tensor(100, device=dev), 1, 1)
ref.optimize_parameters(LRg, HRg, Zg, 2)
full = {k: p.grad.detach().clone() for k, p in ref.G.named_parameters() if p.grad is not None}
assert set(mine) == set(full)
worst = max(rel_l2(mine[k], full[k]) for k in full)
loss_same = abs(float(ref.get_G_train_loss_dict_ref()["xy_gradient"]) - float(gan.get_G_train_loss_dict_ref()["xy_gradient"]))
chk = torch.stack([p.double().sum() for p in gan.G.parameters()]).sum()
dist.all_gather(chks, chk)
res["G step vs concatenated batch"] = (worst, all(bool(c == chks[0]) for c in chks))

# ---- 2. discriminator step (BatchNorm batch statistics and RaGAN batch means stay per rank, stock DDP semantics):
#         the W-rank gradient is the average of the single-rank gradients on the shards
gan, cfg = make(True)
gan.feed_xy_niter(x, y, torch.tensor(100, device=dev), 1, 1)
sdG = {k: v.detach().clone() for k, v in gan.G.state_dict().items()}
sdD = {k: v.detach().clone() for k, v in gan.D.state_dict().items()}
gan.optimize_parameters(LRg[sl], HRg[sl], Zg[sl], 3)
mine = {k: p.grad.detach().clone() for k, p in gan.D.named_parameters() if p.grad is not None}
acc = None
noise = 0.0
for r in range(world):
    grs = []
    for rep in range(2):  # twice: the run-to-run spread of ONE single-rank step is the noise floor of this comparison
        ref, _ = make(False)
        ref.G.load_state_dict(sdG); ref.D.load_state_dict(sdD)
        ref.feed_xy_niter(x, y, torch.tensor(100, device=dev), 1, 1)
        s2 = slice(r * b, (r + 1) * b)
        ref.optimize_parameters(LRg[s2], HRg[s2], Zg[s2], 3)
        grs.append({k: p.grad.detach().clone() for k, p in ref.D.named_parameters() if p.grad is not None})
    sc = max(float(v.double().norm()) for v in grs[0].values())
    noise = max(noise, max(float((grs[0][k].double() - grs[1][k].double()).norm()) /
                           max(float(grs[0][k].double().norm()), 1e-4 * sc) for k in grs[0]))
    gr = grs[0]
    acc = gr if acc is None else {k: acc[k] + gr[k] for k in acc}
assert set(mine) == set(acc)
# (a gradient that is analytically ~0 — the last classifier bias under the symmetric RaGAN loss — has no meaningful
# relative error: judge every tensor against the scale of the whole gradient)
scale = max(float(v.double().norm()) for v in acc.values()) / world
worst = max(float((mine[k].double() - acc[k].double() / world).norm()) / max(float(acc[k].double().norm()) / world, 1e-4 * scale)
            for k in acc)
chk = torch.stack([p.double().sum() for p in gan.D.parameters()]).sum()
dist.all_gather(chks, chk)
res["D step vs average of shards"] = (worst, all(bool(c == chks[0]) for c in chks))
d_noise = noise
if rank == 0:
    print("DDP_RESULT", {k: (float(v[0]), v[1]) for k, v in res.items()}, "D run-to-run noise", d_noise, flush=True)
for name, (worst, same) in res.items():
    # G: every kernel on its path is deterministic in FP32 mode (the concatenated batch only re-associates the batch
    # sums of the weight gradients); D: BatchNorm batch statistics are reduced with fp32 atomics (order varies run to
    # run at the 1e-7 level) and pass through ten normalisations
    # (D: BatchNorm batch statistics are reduced with fp32 atomics; the tiny fixture's batch-of-2 BatchNorm layers
    # amplify that noise — measured 1e-5 .. 9e-3 from run to run — so the bound is the measured run-to-run spread of a
    # single-rank step, with the flat 2e-5 as its floor)
    assert worst <= (5e-6 if name.startswith("G") else max(2e-5, 10.0 * d_noise)), (name, worst, d_noise)
    assert same, name
dist.destroy_process_group()
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_two_rank_step_equals_average_of_single_rank_steps(tmp_path):
    """W = 2 on real GPUs over NCCL: a generator step of the shipped kind (adversarial weight 0) equals the
    single-process step on the CONCATENATED batch (loss normalisers MAX-all-reduced), a discriminator step equals the
    average of the single-rank steps on the shards (per-rank BatchNorm), replicas seeded differently start from rank
    0's weights, and the weights stay identical after each step."""
    script = tmp_path / "ddp_gpu_worker.py"
    script.write_text(_DDP_GPU_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29731", str(script), ROOT],
                       capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "DDP_RESULT" in r.stdout
    print(r.stdout[r.stdout.index("DDP_RESULT"):][:300])
