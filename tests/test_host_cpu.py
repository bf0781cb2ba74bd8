"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol include/windsr.h declares,
the drop-in modules keep the reference's constructors / state_dict keys / seeded initialisation, the ini reader
matches the reference's, the product path refuses to run without CUDA (no fallback), and the data-parallel
gradient exchange averages correctly (gloo, world_size 2)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

from oracle import refshim
from tests.util import GOLDEN, ROOT, load_npz, sd_from, small_generator_kwargs

needs_reference = pytest.mark.skipif(not refshim.available(), reason="reference not mounted")
CONFIGS = os.path.join(GOLDEN, "configs")


def test_library_exports_every_declared_symbol():
    from gan_sr_wind_field_b200 import _lib
    header = open(os.path.join(ROOT, "include", "windsr.h")).read()
    declared = set(re.findall(r"\b(ws_[a-z0-9_]+)\s*\(", header))
    declared -= {"ws_tensor", "ws_conv_shape", "ws_epilogue", "ws_adam_tensor", "ws_prepare_desc", "ws_rdb_desc"}
    bound = {name for name, _, _ in _lib.SYMBOLS}
    assert declared == bound, (declared - bound, bound - declared)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert _lib.load().ws_version() == 1
    # struct layouts agree with the header (sizes the C side was compiled with)
    assert ctypes.sizeof(_lib.WsTensor) == 40 and ctypes.sizeof(_lib.WsConvShape) == 60
    assert ctypes.sizeof(_lib.WsEpilogue) == 3 * 8 + 4 * 4 + 3 * 40 + 16 + 40 + 16 + 2 * 40 + 8  # + tail fields


def test_no_cpu_fallback_and_no_oracle_in_product():
    from gan_sr_wind_field_b200 import _lib, ops
    x = torch.zeros(1, 16, 2, 2, 2)
    w = torch.zeros(8, 16, 3, 3, 3)
    shape = ops.make_shape(x.shape, 8, (3, 3, 3), 1, 1)
    with pytest.raises(_lib.WindSRError):
        ops.conv_fwd(x, w, None, shape, torch.zeros(1, 8, 2, 2, 2))
    with pytest.raises(_lib.WindSRError):
        ops.wind_gradient(torch.zeros(1, 3, 4, 4, 4), torch.arange(4.0), torch.arange(4.0), torch.zeros(1, 1, 4, 4, 4))
    pkg = os.path.join(ROOT, "gan_sr_wind_field_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "cudnn" not in text.lower() or f.endswith(".py") and "no cudnn" in text.lower() or \
                    "cuDNN" in text and "fallback" in text, f


def test_missing_library_fails_loudly(monkeypatch):
    from gan_sr_wind_field_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libwindsr.so")
    with pytest.raises(_lib.WindSRError):
        _lib.load()


def test_state_dict_keys_and_sizes():
    from gan_sr_wind_field_b200.CNN_models.Discriminator_3D import Discriminator_3D
    from gan_sr_wind_field_b200.CNN_models.Generator_3D_Resnet_ESRGAN import Generator_3D
    z = load_npz("generator_small.npz")
    ref_sd = sd_from(z, "sd/")
    G = Generator_3D(**small_generator_kwargs())
    mine = G.state_dict()
    assert list(mine.keys()) == list(ref_sd.keys())
    assert all(mine[k].shape == ref_sd[k].shape for k in mine)
    G.load_state_dict(ref_sd)  # strict
    # the shipped upscale8 architecture: 298 tensors / 35 211 939 params; D: 59 / 12 308 009
    Gf = Generator_3D(4, 3, 128, 16, upscale=8, hr_kern_size=5, lff_kern_size=1, dropout_probability=0.1)
    assert len(Gf.state_dict()) == 298 and sum(p.numel() for p in Gf.parameters()) == 35_211_939
    assert Gf.state_dict()["hr_convs.0.0.weight"].shape == (144, 144, 5, 5, 5)
    assert Gf.state_dict()["model.1.module.3.RDBs.2.conv3.conv.0.weight"].shape == (32, 224, 3, 3, 3)
    Df = Discriminator_3D(3, 32)
    assert len(Df.state_dict()) == 59 and sum(p.numel() for p in Df.parameters()) == 12_308_009
    for tag, slicing in (("full", False), ("slicing", True)):
        ref_d = sd_from(load_npz(f"discriminator_{tag}.npz"), "sd/")
        D = Discriminator_3D(3, 4, enable_slicing=slicing)
        assert list(D.state_dict().keys()) == list(ref_d.keys())
        D.load_state_dict(ref_d)
    # attribute surface callers touch (plot_data.py:773-778, wind_field_GAN_3D.py:581)
    import copy
    assert isinstance(Gf.model[:2], torch.nn.Sequential) and isinstance(Gf.hr_convs[:-2], torch.nn.Sequential)
    feats = copy.deepcopy(Df.features)
    assert sum(p.numel() for p in feats.parameters()) == sum(p.numel() for p in Df.features.parameters())
    assert "Conv3d" in str(Gf) and Gf.max_norm == 1.0


@needs_reference
def test_seeded_construction_equals_reference():
    """Same seed -> bit-identical initial parameters (constructor RNG order + init_weights class-name match)."""
    refshim.activate()
    from CNN_models.Discriminator_3D import Discriminator_3D as RefD
    from CNN_models.Generator_3D_Resnet_ESRGAN import Generator_3D as RefG
    import tools.initialization as ref_init
    from gan_sr_wind_field_b200.CNN_models.Discriminator_3D import Discriminator_3D
    from gan_sr_wind_field_b200.CNN_models.Generator_3D_Resnet_ESRGAN import Generator_3D
    from gan_sr_wind_field_b200.tools import initialization
    kw = small_generator_kwargs()
    torch.manual_seed(5)
    a = RefG(*[kw[k] for k in ("in_channels", "out_channels", "number_of_features", "number_of_RRDBs")],
             **{k: v for k, v in kw.items() if k not in ("in_channels", "out_channels", "number_of_features",
                                                          "number_of_RRDBs")})
    ref_init.init_weights(a, 0.1)
    torch.manual_seed(5)
    b = Generator_3D(**kw)
    initialization.init_weights(b, 0.1)
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa) == list(sb) and all(torch.equal(sa[k], sb[k]) for k in sa)
    for slicing in (False, True):
        torch.manual_seed(6)
        a = RefD(3, 4, enable_slicing=slicing)
        ref_init.init_weights(a, 0.2)
        torch.manual_seed(6)
        b = Discriminator_3D(3, 4, enable_slicing=slicing)
        initialization.init_weights(b, 0.2)
        sa, sb = a.state_dict(), b.state_dict()
        assert list(sa) == list(sb) and all(torch.equal(sa[k], sb[k]) for k in sa)


def test_constructor_errors_match_reference_behaviour():
    from gan_sr_wind_field_b200.CNN_models import torch_blocks as tb
    with pytest.raises(ValueError):
        tb.RDB(16, 8, 5, lff_kern_size=2, mode="3D")
    with pytest.raises(NotImplementedError):
        tb.create_discriminator_block(3, 8, feat_kern_size=7, mode="3D")
    with pytest.raises(NotImplementedError):
        tb.create_conv_lrelu_layer(3, 8, 3, normalization_type="layer")
    with pytest.raises(NotImplementedError):
        tb.RDB(16, 8, 5, mode="weird")


def test_config_reader_on_shipped_inis():
    from gan_sr_wind_field_b200.config.config import Config
    cfg = Config(os.path.join(CONFIGS, "upscale8_pix4_no_adv_no_slicing.ini"))
    assert cfg.scale == 8 and cfg.gpu_id == 0
    g, t = cfg.generator, cfg.training
    assert (g.num_features, g.num_RRDB, g.hr_kern_size, g.lff_kern_size, g.RDB_growth_chan) == (128, 16, 5, 1, 32)
    assert g.conv_mode is None and cfg.gan_config.conv_mode == "3D" and g.dropout_probability == 0.1
    assert t.multistep_lr_steps == [10000, 30000, 50000, 70000, 100000] and t.d_g_train_ratio == 0
    assert (t.pixel_loss_weight, t.gradient_xy_loss_weight, t.gradient_z_loss_weight) == (0.136, 3.064, 0.0)
    assert cfg.dataset_train.batch_size == 8 and cfg.gan_config.enable_slicing is False
    local = Config(os.path.join(CONFIGS, "wind_field_GAN_3D_config_local.ini"))
    assert local.scale == 4 and local.gan_config.enable_slicing is True and local.dataset_train.batch_size == 1
    with pytest.raises(FileNotFoundError):
        Config("/nonexistent.ini")


@needs_reference
def test_config_reader_equals_reference():
    refshim.activate()
    import config.config as ref_config
    from gan_sr_wind_field_b200.config.config import Config
    for name in sorted(os.listdir(CONFIGS)):
        path = os.path.join(CONFIGS, name)
        ref, mine = ref_config.Config(path), Config(path)
        for attr in ("name", "model", "scale", "gpu_id", "use_tensorboard_logger", "load_model_from_save"):
            assert getattr(ref, attr) == getattr(mine, attr), (name, attr)
        for sec in ("env", "gan_config", "generator", "discriminator", "training", "dataset_train", "dataset_val",
                    "dataset_test"):
            r, m = getattr(ref, sec), getattr(mine, sec)
            assert (r is None) == (m is None), (name, sec)
            if r is None:
                continue
            for k, v in vars(r).items():
                assert getattr(m, k) == v, (name, sec, k, v, getattr(m, k))


def test_schedule_and_labels():
    """G/D alternation in blocks of d_g_train_period and the one-sided label smoothing ramp
    (wind_field_GAN_3D.py:585-587, 627-678), decided on the host without touching the device."""
    from gan_sr_wind_field_b200.config.config import Config
    from gan_sr_wind_field_b200.GAN_models.wind_field_GAN_3D import wind_field_GAN_3D
    cfg = Config(os.path.join(CONFIGS, "tiny_gan.ini"))
    cfg.is_train, cfg.device = True, torch.device("cpu")
    gan = wind_field_GAN_3D(cfg)
    gan.feed_xy_niter(torch.arange(4.0), torch.arange(4.0), torch.tensor(100), 1, 2)
    assert [gan.is_G_iteration(i) for i in range(8)] == [True, True, False, False, True, True, False, False]
    gan.batch_size = 3
    gan.make_new_labels(0)
    assert torch.allclose(gan.HR_labels, torch.full((3,), 0.9)) and float(gan._labels_point_nine) == 1.0
    gan.make_new_labels(50)
    assert torch.allclose(gan.HR_labels, torch.full((3,), 0.95)) and float(gan._labels_point_nine) == 0.0
    # instance-noise scales sqrt(sigma * (1 - (it-1)/niter)) for sigma = 1, 2 (trainingtricks.py:49-58)
    assert torch.allclose(gan._scalars[2:4], torch.tensor([0.51 ** 0.5, 1.02 ** 0.5]))
    assert torch.equal(gan.fake_HR_labels, torch.zeros(3))
    assert gan.count_params() == (sum(p.numel() for p in gan.G.parameters()),
                                  sum(p.numel() for p in gan.D.parameters()))


_DDP_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from gan_sr_wind_field_b200.parallel import GradSync, allreduce_max, allreduce_max_, broadcast_module
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
torch.manual_seed(0)
net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Tanh(), torch.nn.Linear(16, 4), torch.nn.Linear(4, 1))
frozen = net[2].bias
frozen.requires_grad = False                      # like D's parameters during a G step
sync = GradSync(net.parameters(), bucket_bytes=256)   # several small buckets
data = torch.randn(world, 5, 8, generator=torch.Generator().manual_seed(1))
sync.begin()
net(data[rank]).pow(2).mean().backward()
sync.finish()
# reference: mean over ranks of the per-rank gradients, computed locally
ref = [torch.zeros_like(p) for p in net.parameters()]
for r in range(world):
    net2 = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Tanh(), torch.nn.Linear(16, 4), torch.nn.Linear(4, 1))
    net2.load_state_dict(net.state_dict())
    net2[2].bias.requires_grad = False
    net2(data[r]).pow(2).mean().backward()
    for acc, p in zip(ref, net2.parameters()):
        if p.grad is not None:
            acc += p.grad / world
ok = all((p.grad is None and not p.requires_grad) or torch.allclose(p.grad, r_, atol=1e-6)
         for p, r_ in zip(net.parameters(), ref))
# a second step re-arms cleanly
net.zero_grad(set_to_none=True)
sync.begin(); net(data[rank]).pow(2).mean().backward(); sync.finish()
ok = ok and all((p.grad is None) or torch.allclose(p.grad, r_, atol=1e-6) for p, r_ in zip(net.parameters(), ref))
# p.grad is a view of the reduced bucket (no copy back)
live = [p for p in net.parameters() if p.grad is not None]
def flat_ptr(p):
    b = sync._where[p]
    i = [id(q) for q in b.params].index(id(p))
    return b.flat[b.offsets[i]:].data_ptr()
ok = ok and all(p.grad.data_ptr() == flat_ptr(p) for p in live)
# replicas built from different seeds take rank 0's parameters
torch.manual_seed(10 + rank)
other = torch.nn.Linear(3, 3)
broadcast_module(other)
gathered = [torch.zeros_like(other.weight) for _ in range(world)]
dist.all_gather(gathered, other.weight.detach())
ok = ok and all(torch.equal(g, gathered[0]) for g in gathered)
# the global "loss is not finite" flag: one bad rank makes every rank skip
flag = torch.tensor(1.0 if rank == 1 else 0.0)
ok = ok and float(allreduce_max_(flag)) == 1.0
# differentiable MAX over ranks (the loss normalisers): forward = global max; backward = the SUM of all ranks'
# cotangents, delivered to the rank that holds the maximum only
x = torch.tensor([1.0 + rank, 5.0 - rank], requires_grad=True)     # slot 0: rank 1 wins, slot 1: rank 0 wins
y = allreduce_max(x)
ok = ok and torch.equal(y.detach(), torch.tensor([2.0, 5.0]))
(y * torch.tensor([1.0 + rank, 10.0 * (1 + rank)])).sum().backward()  # cotangents (1|2, 10|20) on rank (0|1)
want = torch.tensor([0.0, 30.0]) if rank == 0 else torch.tensor([3.0, 0.0])
ok = ok and torch.equal(x.grad, want)
dist.barrier(); dist.destroy_process_group()
sys.exit(0 if ok else 3)
'''


def test_grad_sync_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_DDP_WORKER)
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29641")
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT], env=env))
    codes = [p.wait(timeout=180) for p in procs]
    assert codes == [0, 0], codes


def test_optimizer_step_invalidates_packed_weights():
    """Every optimizer step (fused ones do not bump Tensor._version) must advance the packed-weight epoch."""
    from gan_sr_wind_field_b200 import ops
    p = torch.nn.Parameter(torch.randn(4, 4))
    p.grad = torch.randn(4, 4)
    e0 = ops._WEIGHTS_EPOCH
    torch.optim.Adam([p], fused=True).step()
    assert ops._WEIGHTS_EPOCH == e0 + 1
    torch.optim.SGD([p], lr=0.1).step()
    assert ops._WEIGHTS_EPOCH == e0 + 2


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the reference's CPU path: the stock modules from baseline/_ref when installed, else
    the oracle port; no GPU needed) prints ONE JSON line with the keys the driver reads."""
    import json
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--ref-batch", "1"], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "GAN train voxels/sec" and line["unit"] == "HR voxels/s"
    for k in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] in ("port", "reference") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"]


def test_im2col_weight_reordering_matches_conv():
    """Host-side transform of the narrow first layers: conv(x, w) == U @ w'^T with U[v, tap*cin+ci] = x[ci, v (+) tap]
    (what ws_im2col builds on the device) and w' = ops._im2col_weight(w) — checked with torch ops on the CPU."""
    import torch.nn.functional as F
    from gan_sr_wind_field_b200 import ops
    torch.manual_seed(3)
    cin, cout, k = 3, 32, 3
    x = torch.randn(2, cin, 6, 5, 4)
    w = torch.randn(cout, cin, k, k, k)
    ref = F.conv3d(x, w, padding=1)
    xp = F.pad(x, (1, 1, 1, 1, 1, 1))
    cols = []
    for ti in range(k):
        for tj in range(k):
            for tl in range(k):
                for ci in range(cin):
                    cols.append(xp[:, ci, ti:ti + 6, tj:tj + 5, tl:tl + 4])
    U = torch.stack(cols, 1)                      # (n, taps*cin, X, Y, Z), tap-major / ci-minor columns
    cpad = 96
    w2 = ops._im2col_weight(w, cpad).reshape(cout, cpad)
    assert torch.all(w2[:, k ** 3 * cin:] == 0)
    out = torch.einsum("ncxyz,oc->noxyz", U, w2[:, :k ** 3 * cin])
    assert torch.allclose(out, ref, atol=1e-4, rtol=1e-4)


def test_xfold_weight_matches_conv():
    """x-fold of a narrow-output conv (hr_convs.2): a (1,ky,kz) conv with the kx taps stacked on the output channels
    followed by y[x] = sum_dx U[x + dx - px][dx] equals the direct conv — checked with torch ops on the CPU."""
    import torch.nn.functional as F
    from gan_sr_wind_field_b200 import ops
    torch.manual_seed(4)
    co, ci, k, p = 3, 8, 5, 2
    x = torch.randn(1, ci, 7, 6, 5)
    w = torch.randn(co, ci, k, k, k)
    ref = F.conv3d(x, w, padding=p)
    w5 = ops.XFoldConvFn._folded_weight(w)        # (16, ci, 1, k, k): rows dx*co + c, zero padded to 16
    assert w5.shape == (16, ci, 1, k, k) and torch.all(w5[k * co:] == 0)
    U = F.conv3d(x, w5, padding=(0, p, p))        # (1, 16, X, Y, Z)
    X = x.shape[2]
    out = torch.zeros_like(ref)
    for dx in range(k):
        for xx in range(X):
            xs = xx + dx - p
            if 0 <= xs < X:
                out[:, :, xx] += U[:, dx * co:(dx + 1) * co, xs]
    assert torch.allclose(out, ref, atol=1e-4, rtol=1e-4)


def test_device_dataset_sampling_follows_reference_distributions(monkeypatch):
    """Host side of the GPU input pipeline: per-sample draws in the reference's order and ranges
    (process_data.py:159-166, 199, 246, 252); no device work (tensors stay on the CPU here)."""
    import numpy as np
    from gan_sr_wind_field_b200.data_pipeline import DeviceWindDataset
    f = np.zeros((3, 24, 20, 4))
    ds = DeviceWindDataset(f, f, f, f, f, f, -2.71, 550.44, 32.33, 9e4, 1.05e5, 68.46, include_z_channel=True,
                           COARSENESS_FACTOR=4, enable_slicing=True, slice_size=16, device="cpu", seed=7)
    aug = ds.sample_augmentation(2000)
    assert aug.dtype == np.int32 and aug.shape == (2000, 5)
    assert aug[:, 0].min() >= 0 and aug[:, 0].max() <= 24 - 16 and aug[:, 1].max() <= 20 - 16
    # beta(0.25, 0.25) piles the offsets up at both ends of the range
    assert (aug[:, 0] == 0).mean() > 0.2 and (aug[:, 0] == 8).mean() > 0.2
    assert set(np.unique(aug[:, 2])) == {0, 1, 2, 3} and set(np.unique(aug[:, 3:])) == {0, 1}
    plain = DeviceWindDataset(f, f, f, f, f, f, -2.71, 550.44, 32.33, 9e4, 1.05e5, 68.46, data_aug_rot=False,
                              data_aug_flip=False, device="cpu")
    assert not plain.sample_augmentation(5).any() and len(plain) == 3


def test_trunk_batching_host_logic():
    """Host side of the batched trunk path (ops.TrunkFn / ws_trunk_wgrad): which runs of RRDBs are recognised as
    batchable, how the blocks' epilogue constants and outer-skip links are laid out, and that CPU tensors never take it
    (torch_blocks.py:217-330 of the reference define the structure being matched)."""
    import torch
    from gan_sr_wind_field_b200 import ops
    from gan_sr_wind_field_b200.CNN_models import torch_blocks as tb
    from gan_sr_wind_field_b200.CNN_models.Generator_3D_Resnet_ESRGAN import Generator_3D
    G = Generator_3D(4, 3, 128, 2, upscale=8, hr_kern_size=5, number_of_RDB_convs=5, RDB_gc=32, lff_kern_size=1,
                     terrain_number_of_features=16)
    rrdbs = [m for m in G.model[1].module if isinstance(m, tb.RRDB)]
    assert len(rrdbs) == 2
    sig = tb._trunk_signature(rrdbs[0])
    assert sig is not None and sig == tb._trunk_signature(rrdbs[1])
    assert sig[:4] == (128, 32, 4, 3)                       # features, growth channels, dense convs, kernel size
    assert tb._trunk_signature(G.model[1].module[-1]) is None  # lr_conv is not an RRDB
    odd = tb.RRDB(128, 32, 5, 3, mode="3D")                  # 3x3x3 LFF: not the batched 1x1x1 case
    assert tb._trunk_signature(odd) is None
    other = tb.RRDB(128, 16, 5, 1, mode="3D")                # different growth channels: a different signature
    assert tb._trunk_signature(other) not in (None, sig)
    # block table of one RRDB: RDB 0/1 plain, the last one carries the RRDB scale and the outer skip to block 0's input
    captured = {}

    class _Stop(Exception):
        pass

    def fake_apply(x, cfg, *params):
        captured["cfg"], captured["n"] = cfg, len(params)
        raise _Stop()

    real = ops.TrunkFn.apply
    ops.TrunkFn.apply = staticmethod(fake_apply)
    try:
        try:
            tb._trunk_apply(rrdbs, torch.zeros(1, 128, 4, 4, 2), sig)
        except _Stop:
            pass
    finally:
        ops.TrunkFn.apply = real
    blocks = captured["cfg"]["blocks"]
    assert len(blocks) == 6 and captured["n"] == 6 * 6     # per block: 4 dense weights, LFF weight, LFF bias
    assert [b["outer"] for b in blocks] == [None, None, 0, None, None, 3]
    assert abs(blocks[2]["alpha"] - 0.2 * 0.2) < 1e-12 and blocks[2]["beta1"] == 0.2 and blocks[2]["beta2"] == 1.0
    assert blocks[0]["alpha"] == 0.2 and blocks[0]["beta1"] == 1.0 and blocks[0]["beta2"] == 0.0
    # a CPU tensor goes through the modules' own forward, which refuses (no CPU fallback) instead of batching
    with pytest.raises(Exception):
        tb.run_trunk(rrdbs, torch.zeros(1, 128, 4, 4, 2))
