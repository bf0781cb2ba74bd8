"""``wind_field_GAN_3D``: the GAN training step that drives the hot path.

Public surface, loss definitions and schedule follow GAN_models/wind_field_GAN_3D.py of the reference
(constructor :30-205, ``feed_xy_niter`` :207-219, ``optimize_parameters`` / ``validation`` :621-625, loss
dicts :680-693, ``update_learning_rate`` :695-697, ``count_params`` :699-712).  What differs is *how* the step
is executed:

* G and D are the drop-in modules running on ``libwindsr.so``;
* the pixel loss, both 9-channel Jacobians, the 8 global maxes and the 4 MSE sums are ONE fused stencil pass
  (``ops.WindLossFn``); the scalar formula on top is a handful of 0-d tensor ops;
* the schedule is decided from the host-side iteration counter and the 8 ``isnan/isinf`` probes collapse to a
  single host read per G step (the reference syncs >= 9 times per step);
* label noise is sampled on the device; every iteration-dependent scalar (label values, instance-noise scales, the
  "labels are exactly 0.9" switch, learning rate) lives in a small device vector, so that
* a whole G step and a whole D step are each captured ONCE into a CUDA graph (after a few eager warm-up calls) and
  replayed — ~2 000 kernel launches per step cost the host one ``cudaGraphLaunch`` (``WINDSR_CUDA_GRAPH=0`` keeps
  the eager path);
* Adam is ``optim.WindAdam`` (one hand-written multi-tensor kernel with a device-side skip flag);
* with ``adversarial_loss_weight == 0`` (both shipped pretrained configs) a G step does not run the discriminator:
  the reference multiplies that branch by 0.0 (:426), so loss and gradients are unchanged
  (``WINDSR_SKIP_D_WHEN_ZERO=0`` runs it anyway, e.g. to reproduce a NaN coming out of D);
* with ``torch.distributed`` initialised the step is batch-sharded data parallel: gradients of the network
  being updated are averaged with a bucketed NCCL all-reduce overlapped with backward (``parallel.GradSync``).
  BatchNorm statistics, RaGAN batch means and the loss normalisers stay per-rank (stock DDP semantics; see
  DESIGN.md §multi-GPU for the exactness statement).
"""
from __future__ import annotations

import copy
import os
import math

import torch
import torch.nn as nn
import torch.nn.functional as F
import torch.optim.lr_scheduler as lr_scheduler

from .. import ops
from ..CNN_models.Discriminator_3D import Discriminator_3D
from ..CNN_models.Generator_3D_Resnet_ESRGAN import Generator_3D
from ..optim import WindAdam
from ..parallel import GradSync, allreduce_max, allreduce_max_, broadcast_module
from ..tools import initialization, trainingtricks
from .baseGAN import BaseGAN

_G_LOSS_KEYS = ("total", "adversarial", "pix", "xy_gradient", "z_gradient", "divergence", "xy_divergence",
                "feature_D")


def _zero():
    return torch.zeros(1)


def _leaf_probe(module):
    """The LAST leaf sub-module: the cheapest place where a partial ``.eval()`` / ``.train()`` done behind this
    class's back (``D.features.eval()``) shows up next to the top-level flag."""
    probe = getattr(module, "_windsr_mode_probe", None)
    if probe is None:
        leaves = [m for m in module.modules() if not list(m.children())]
        probe = (leaves[0], leaves[len(leaves) // 2], leaves[-1]) if leaves else (module,)
        object.__setattr__(module, "_windsr_mode_probe", probe)
    return probe


def _set_mode(module, training: bool):
    """module.train()/.eval() walks every sub-module (~1000 here): skip the walk when the top-level flag AND three
    probe leaves (first / middle / last) already agree; the reference re-applies the mode every step
    (wind_field_GAN_3D.py:248,274,478), so a mode changed by outside code must not survive."""
    if module.training != training or any(m.training != training for m in _leaf_probe(module)):
        module.train(training)


def _set_requires_grad(module, flag: bool):
    """``for p in D.parameters(): p.requires_grad = flag`` (wind_field_GAN_3D.py:481-482,536-537) without the walk
    when first / last parameters already agree."""
    params = getattr(module, "_windsr_param_probe", None)
    if params is None:
        ps = list(module.parameters())
        params = (ps[0], ps[len(ps) // 2], ps[-1]) if ps else ()
        object.__setattr__(module, "_windsr_param_probe", params)
    if any(p.requires_grad != flag for p in params):
        for p in module.parameters():
            p.requires_grad = flag


class wind_field_GAN_3D(BaseGAN):
    def __init__(self, cfg):
        super().__init__(cfg)
        self.train_G_loss_dict = {k: _zero() for k in _G_LOSS_KEYS}
        self.validation_G_loss_dict = {k: _zero() for k in _G_LOSS_KEYS}
        self.D_loss_dict = {"train_loss": _zero(), "validation_loss": _zero()}
        self.hist_dict = {k: _zero() for k in ("val_grad_G_first_layer", "val_grad_G_last_layer",
                                               "val_weight_G_first_layer", "val_weight_G_last_layer",
                                               "SR_pix_distribution", "D_pred_HR", "D_pred_SR")}
        for k in ("val_grad_D_first_layer", "val_grad_D_last_layer", "val_weight_D_first_layer",
                  "val_weight_D_last_layer"):
            self.hist_dict[k] = torch.tensor(-1.0)
        self.metrics_dict = {k: _zero() for k in ("val_PSNR", "Trilinear_PSNR", "pix_loss_unscaled",
                                                  "trilinear_pix_loss")}
        self.batch_size = 1
        self.max_diff_squared = torch.tensor(4.0, device=self.device)  # HR is in [-1, 1]
        self.epsilon_PSNR = torch.tensor(1e-8, device=self.device)
        self.feature_extractor = None
        self._inflight = []  # CUDA events of the last training steps (bounded host run-ahead)
        self._graphs = {}       # (kind, shapes, precision) -> graph_step.StepGraph
        self._eager_calls = {}  # the same key -> eager calls so far (a step is captured after a few of them)
        self._scalars = None    # device vector of the iteration-dependent scalars (see _write_scalars)
        self._scalar_ring, self._scalar_slot = [], 0
        self.rank = torch.distributed.get_rank() if torch.distributed.is_initialized() else 0
        self.world_size = torch.distributed.get_world_size() if torch.distributed.is_initialized() else 1

        g, gan = cfg.generator, cfg.gan_config
        in_ch = (g.in_num_ch + int(bool(gan.include_pressure)) + int(bool(gan.include_z_channel))
                 + int(bool(gan.include_above_ground_channel)))
        self.G = Generator_3D(
            in_ch, g.out_num_ch, g.num_features, g.num_RRDB, upscale=cfg.scale, hr_kern_size=g.hr_kern_size,
            number_of_RDB_convs=g.num_RDB_convs, RDB_gc=g.RDB_growth_chan, lff_kern_size=g.lff_kern_size,
            RDB_residual_scaling=g.RDB_res_scaling, RRDB_residual_scaling=g.RRDB_res_scaling,
            act_type=g.act_type, device=self.device, number_of_z_layers=gan.number_of_z_layers,
            conv_mode=gan.conv_mode, use_mixed_precision=g.use_mixed_precision,
            terrain_number_of_features=g.terrain_number_of_features,
            dropout_probability=g.dropout_probability, max_norm=g.max_norm,
        ).to(self.device, non_blocking=True)
        initialization.init_weights(self.G, scale=g.weight_init_scale)
        self.conv_mode = g.conv_mode
        self.use_D_feature_extractor_cost = gan.use_D_feature_extractor_cost
        self.D = None
        self.sync_G = self.sync_D = None
        if not cfg.is_train:
            return

        d = cfg.discriminator
        self.D = Discriminator_3D(
            d.in_num_ch, d.num_features, feat_kern_size=d.feat_kern_size, normalization_type=d.norm_type,
            act_type=d.act_type, mode=d.layer_mode, device=self.device,
            number_of_z_layers=gan.number_of_z_layers, conv_mode=gan.conv_mode,
            use_mixed_precision=d.use_mixed_precision, enable_slicing=gan.enable_slicing,
            dropout_probability=d.dropout_probability,
        ).to(self.device, non_blocking=True)
        initialization.init_weights(self.D, scale=d.weight_init_scale)

        t = cfg.training
        if self.world_size > 1:
            # replicas must start from the same weights / BatchNorm buffers whatever the callers seeded, and then
            # draw DIFFERENT dropout / label-noise / instance-noise streams
            broadcast_module(self.G)
            broadcast_module(self.D)
            seed = int(getattr(getattr(cfg, "env", None), "fixed_seed", 0) or 0)
            torch.manual_seed(seed + self.rank)
        # Same Adam as wind_field_GAN_3D.py:147-160, as one hand-written multi-tensor kernel with a device-side
        # ``found_inf`` flag: the reference's "skip the step when the loss is not finite" guard (:457) without a
        # host read.  WINDSR_FUSED_ADAM=0: torch's own (host-checked guard).
        self._fused_adam = self.device.type == "cuda" and os.environ.get("WINDSR_FUSED_ADAM", "1") != "0"
        adam = WindAdam if self._fused_adam else torch.optim.Adam
        self.optimizer_G = adam(self.G.parameters(), lr=t.learning_rate_g, weight_decay=t.adam_weight_decay_g,
                                betas=(t.adam_beta1_g, 0.999))
        self.optimizer_D = adam(self.D.parameters(), lr=t.learning_rate_d, weight_decay=t.adam_weight_decay_d,
                                betas=(t.adam_beta1_d, 0.999))
        self.optimizers += [self.optimizer_G, self.optimizer_D]
        if t.multistep_lr_steps:
            self.scheduler_G = lr_scheduler.MultiStepLR(self.optimizer_G, t.multistep_lr_steps, gamma=t.lr_gamma)
            self.scheduler_D = lr_scheduler.MultiStepLR(self.optimizer_D, t.multistep_lr_steps, gamma=t.lr_gamma)
            self.schedulers += [self.scheduler_G, self.scheduler_D]
        if t.pixel_criterion in (None, "none"):
            self.pixel_criterion = None
        elif t.pixel_criterion in ("l1", "l2"):
            self.pixel_criterion = t.pixel_criterion
        else:
            raise NotImplementedError("Only l1 and l2 (MSE) loss have been implemented for pixel loss, not "
                                      f"{t.pixel_criterion}")
        if t.gan_type not in ("relativistic", "relativisticavg"):
            raise NotImplementedError(f"Only relativistic and relativisticavg GAN are implemented, not {t.gan_type}")
        self.criterion = nn.BCEWithLogitsLoss()
        if self.world_size > 1:
            bucket = int(float(os.environ.get("WINDSR_BUCKET_MB", "25")) * (1 << 20))
            self.sync_G = GradSync(self.G.parameters(), bucket_bytes=bucket)
            self.sync_D = GradSync(self.D.parameters(), bucket_bytes=bucket)

    # ---------------------------------------------------------------------------------------------------
    def feed_xy_niter(self, x, y, niter, d_g_train_ratio, d_g_train_period):
        self.x, self.y = x, y
        self.niter = niter
        self._niter_host = int(niter)
        self.d_g_train_ratio = d_g_train_ratio
        self.d_g_train_period = d_g_train_period

    # ---------------------------------------------------------------------------------------------------
    # iteration-dependent scalars, kept on the device (so a captured step can be replayed for any iteration)
    _S_REAL, _S_FAKE, _S_NOISE1, _S_NOISE2, _S_POINT_NINE, _S_COUNT = 0, 1, 2, 3, 4, 8

    def _write_scalars(self, it: int):
        """real / fake label values (wind_field_GAN_3D.py:627-678), the instance-noise scales
        sqrt(sigma * (1 - (it-1)/niter)) for sigma = 1 (D steps) and 2 (G steps) (trainingtricks.py:49-58; NaN
        past niter+1 like the reference's sqrt of a negative), and the "real labels are exactly 0.9" switch of
        :557-558 — computed on the host, sent with ONE small asynchronous copy from a ring of pinned buffers."""
        t = self.cfg.training
        real, fake = 1.0, 0.0
        frac = float(it) / float(self._niter_host)
        if t.use_one_sided_label_smoothing and t.flip_labels:
            fake = 0.1 - 0.1 * frac
        elif t.use_one_sided_label_smoothing:
            real = 0.9 + 0.1 * frac
        var = 1.0 - (float(it) - 1.0) / float(self._niter_host)
        n1 = math.sqrt(var) if var >= 0.0 else float("nan")
        n2 = math.sqrt(2.0 * var) if var >= 0.0 else float("nan")
        target = real if not t.flip_labels else fake
        # float32(0.9 + 0.1*frac) == float32(0.9) exactly as the reference's device-side comparison sees it
        nine = float(torch.tensor(target, dtype=torch.float32) == torch.tensor(0.9, dtype=torch.float32))
        vals = [real, fake, n1, n2, nine, 0.0, 0.0, 0.0]
        if self._scalars is None:
            self._scalars = torch.zeros(self._S_COUNT, dtype=torch.float32, device=self.device)
            if self.device.type == "cuda":
                self._scalar_ring = [torch.zeros(self._S_COUNT, dtype=torch.float32).pin_memory() for _ in range(8)]
        if self._scalar_ring:
            # the host never runs more than two steps ahead (optimize_parameters), so 8 slots cannot be overwritten
            # before their copy has executed
            h = self._scalar_ring[self._scalar_slot]
            self._scalar_slot = (self._scalar_slot + 1) % len(self._scalar_ring)
            h.copy_(torch.tensor(vals, dtype=torch.float32))
            self._scalars.copy_(h, non_blocking=True)
        else:
            self._scalars.copy_(torch.tensor(vals, dtype=torch.float32))

    def _noisy(self, x, which: int):
        """x + instance noise (wind_field_GAN_3D.py:250-299): one fused kernel on the GPU."""
        scale = self._scalars[which:which + 1]
        if x.is_cuda:
            return ops.add_instance_noise(x, 1.0, scale_dev=scale)
        return x + torch.rand(x.shape, device=x.device) * scale

    def D_forward(self, HR, fake_HR, it, train_D: bool):
        """D on the real and generated batch (wind_field_GAN_3D.py:221-304): train mode + sigma 1 noise in D
        steps; eval mode (BN running stats, no dropout) + sigma 2 noise, real branch detached, in G steps."""
        noisy = self.cfg.training.use_instance_noise
        if train_D:
            _set_mode(self.D, True)
            y_pred = self.D(self._noisy(HR, self._S_NOISE1) if noisy else HR).squeeze()
            fake_in = fake_HR.detach()
            return y_pred, self.D(self._noisy(fake_in, self._S_NOISE1) if noisy else fake_in).squeeze()
        _set_mode(self.D, False)
        with torch.no_grad():
            y_pred = self.D(self._noisy(HR, self._S_NOISE2) if noisy else HR).squeeze()
        return y_pred, self.D(self._noisy(fake_HR, self._S_NOISE2) if noisy else fake_HR).squeeze()

    # ---------------------------------------------------------------------------------------------------
    def _adversarial(self, first, second, generator_side: bool):
        """relativistic / relativistic-average BCE terms (wind_field_GAN_3D.py:353-368, 545-563)."""
        kind = self.cfg.training.gan_type
        if kind == "relativistic":
            return self.criterion(first - second, self.HR_labels)
        return (self.criterion(first - torch.mean(second), self.HR_labels)
                + self.criterion(second - torch.mean(first), self.fake_HR_labels)) / 2.0

    def wind_loss_terms(self, HR, fake_HR, Z):
        """pixel + the four physics terms, unweighted, from one fused stencil pass.
        Returns (pix, xy_gradient, z_gradient, divergence, xy_divergence)."""
        if HR.shape[1] != 3 or fake_HR.shape[1] != 3:
            raise NotImplementedError("the fused wind loss expects exactly the 3 wind components (all shipped "
                                      "configs: out_num_ch = 3)")
        s = ops.windloss_slots(HR, fake_HR, Z, self.x, self.y)
        cnt = float(HR.shape[0] * HR.shape[2] * HR.shape[3] * HR.shape[4])
        mx = s[6:14]
        if self.world_size > 1 and os.environ.get("WINDSR_SYNC_NORMALISERS", "1") != "0":
            # the four loss normalisers are maxima over the WHOLE batch (wind_field_GAN_3D.py:773-814): with the batch
            # sharded over ranks they are the one batch-coupled term of a generator step (G has no BatchNorm), so one
            # 8-float MAX all-reduce makes W ranks reproduce the reference on the concatenated batch
            mx = allreduce_max(mx)
        norm = [torch.maximum(mx[2 * k], mx[2 * k + 1] / 100) for k in range(4)]
        xy = s[0] / (6.0 * cnt) / (norm[0] * norm[0])
        zg = s[1] / (3.0 * cnt) / (norm[1] * norm[1])
        div = s[2] / cnt / (norm[2] * norm[2])
        dxy = s[3] / cnt / (norm[3] * norm[3])
        if self.pixel_criterion == "l1":
            pix = s[4] / (3.0 * cnt)
        elif self.pixel_criterion == "l2":
            pix = s[5] / (3.0 * cnt)
        else:
            pix = torch.zeros((), device=self.device)
        return pix, xy, zg, div, dxy

    def _skip_D_in_G_step(self) -> bool:
        """adversarial weight exactly 0 and no feature-extractor term: the discriminator contributes exact zeros to
        the generator loss and to its gradients (the reference still runs it, wind_field_GAN_3D.py:487,426)."""
        return (float(self.cfg.training.adversarial_loss_weight) == 0.0 and self.feature_extractor is None
                and not self.use_D_feature_extractor_cost and os.environ.get("WINDSR_SKIP_D_WHEN_ZERO", "1") != "0")

    def calculate_optimize_and_log_G_loss(self, HR, fake_HR, Z, y_pred, fake_y_pred, training_iteration: bool):
        t = self.cfg.training
        if y_pred is None:
            adv = torch.zeros((), device=self.device)
        else:
            adv = self._adversarial(fake_y_pred, y_pred, True) * t.adversarial_loss_weight
        feat = torch.zeros(1, device=self.device)
        if self.feature_extractor is not None:
            with torch.no_grad():
                f_real = self.feature_extractor(HR)
            feat = F.mse_loss(self.feature_extractor(fake_HR).float(), f_real.float())
        feat = feat * t.feature_D_loss_weight
        pix, xy, zg, div, dxy = self.wind_loss_terms(HR, fake_HR, Z)
        pix = pix * t.pixel_loss_weight
        xy = xy * t.gradient_xy_loss_weight
        zg = zg * t.gradient_z_loss_weight
        div = div * t.divergence_loss_weight
        dxy = dxy * t.xy_divergence_loss_weight
        base = adv + pix + feat
        physics = xy + zg + div + dxy
        # The reference's guards (:434-443 and :457): drop the physics terms when any of them is NaN/Inf, and skip
        # the optimiser step when the total is NaN/Inf.  The first is decided ON THE DEVICE (select, with
        # WindLossFn.backward discarding the 0*inf cotangents of the dropped branch); the second is handed to the
        # Adam kernel as its ``found_inf`` flag, so a training step has no host synchronisation at all.
        ok_physics = torch.isfinite(physics).all()
        loss_G = base + torch.where(ok_physics, physics, torch.zeros_like(physics))
        if training_iteration:
            if self.sync_G is not None:
                self.sync_G.begin()
            loss_G.backward()
            # data parallel: the gradients are averaged over ranks, so the skip decision must be global too —
            # one rank's NaN poisons everybody's average and every replica has to skip the same step
            not_finite = (~torch.isfinite(loss_G.detach()).reshape(())).to(torch.float32)
            if self.sync_G is not None:
                allreduce_max_(not_finite)
                self.sync_G.finish()
            if self._fused_adam:
                self.optimizer_G.found_inf = not_finite
                self.optimizer_G.step()  # a no-op on the device when loss_G is not finite on any rank
            elif not bool(not_finite):
                self.optimizer_G.step()
        d = self.train_G_loss_dict if training_iteration else self.validation_G_loss_dict
        # detached: a stored loss with its grad_fn would keep this step's whole autograd graph (and the parameters'
        # AccumulateGrad nodes, bound to the stream they were created on) alive into the next step — ~6 GB of saved
        # activations, and a stream mismatch that invalidates a CUDA-graph capture of the next step
        det = lambda t: t.detach()
        d.update(total=det(loss_G), adversarial=det(adv), pix=det(pix), xy_gradient=det(xy), z_gradient=det(zg),
                 divergence=det(div), xy_divergence=det(dxy), feature_D=det(feat))
        if not training_iteration:
            self.metrics_dict["pix_loss_unscaled"] = pix.detach() / t.pixel_loss_weight
            self.hist_dict["SR_pix_distribution"] = fake_HR.detach().cpu().numpy()
        return loss_G

    def update_G(self, LR, HR, Z, it, training_iteration: bool):
        if training_iteration:
            _set_mode(self.G, True)
            fake_HR = self.G(LR, Z)
            _set_requires_grad(self.D, False)
            self.G.zero_grad(set_to_none=True)
            if self._skip_D_in_G_step():
                y_pred = fake_y_pred = None
            else:
                y_pred, fake_y_pred = self.D_forward(HR, fake_HR, it, train_D=False)
            self.calculate_optimize_and_log_G_loss(HR, fake_HR, Z, y_pred, fake_y_pred, True)
        else:
            _set_mode(self.G, False)
            with torch.no_grad():
                fake_HR = self.G(LR, Z)
                y_pred, fake_y_pred = self.D_forward(HR, fake_HR, it, train_D=False)
                self.calculate_optimize_and_log_G_loss(HR, fake_HR, Z, y_pred, fake_y_pred, False)
        return fake_HR

    def update_D(self, HR, fake_HR, it, training_epoch: bool):
        if training_epoch:
            _set_requires_grad(self.D, True)
            self.optimizer_D.zero_grad(set_to_none=True)
            y_pred, fake_y_pred = self.D_forward(HR, fake_HR, it, train_D=True)
        else:
            with torch.no_grad():  # the reference validates D in train mode too (:542-543)
                y_pred, fake_y_pred = self.D_forward(HR, fake_HR, it, train_D=True)
        loss_D = self._adversarial(y_pred, fake_y_pred, False)
        if self.cfg.training.gan_type == "relativisticavg":
            # `if torch.all(self.HR_labels == 0.9): loss_D -= 0.1985` (:557-558) as a device-side select
            loss_D = loss_D - 0.1985 * self._labels_point_nine
        if training_epoch:
            if self.sync_D is not None:
                self.sync_D.begin()
            loss_D.backward()
            if self.sync_D is not None:
                self.sync_D.finish()
            if self._fused_adam:
                self.optimizer_D.found_inf = None
            self.optimizer_D.step()
        if training_epoch:
            self.D_loss_dict["train_loss"] = loss_D.detach()
        else:
            self.D_loss_dict["validation_loss"] = loss_D.detach()
            self.hist_dict["D_pred_HR"] = torch.sigmoid(y_pred.detach()).cpu().numpy()[None]
            self.hist_dict["D_pred_SR"] = torch.sigmoid(fake_y_pred.detach()).cpu().numpy()[None]

    # ---------------------------------------------------------------------------------------------------
    def is_G_iteration(self, it: int) -> bool:
        """Alternating blocks of ``d_g_train_period`` iterations (wind_field_GAN_3D.py:585-587)."""
        return (int(it) // self.d_g_train_period) % (self.d_g_train_ratio + 1) == 0

    def _train_step_body(self, kind: str, LR, HR, Z):
        """Everything of one training iteration that runs on the device; depends on the iteration number only
        through ``self._scalars`` — this is what gets captured into a CUDA graph."""
        self.make_new_labels()
        if kind == "G":
            self.update_G(LR, HR, Z, None, True)
        else:
            with torch.no_grad():
                _set_mode(self.G, False)
                fake_HR = self.G(LR, Z)
            self.update_D(HR, fake_HR, None, True)

    def _graph_eligible(self, LR) -> bool:
        return (LR.is_cuda and self._fused_adam and not self.use_D_feature_extractor_cost
                and os.environ.get("WINDSR_CUDA_GRAPH", "1") != "0")

    def compute_losses_and_optimize(self, LR, HR, Z, it, training_iteration: bool = False):
        self.batch_size = HR.size(0)
        it_host = int(it)
        self._write_scalars(it_host)
        if self.use_D_feature_extractor_cost and it_host % self.cfg.training.feature_D_update_period == 0:
            self.feature_extractor = copy.deepcopy(self.D.features)
            for p in self.feature_extractor.parameters():
                p.requires_grad = False
        if training_iteration:
            kind = "G" if self.is_G_iteration(it_host) else "D"
            if self._graph_eligible(LR):
                from .graph_step import run_captured
                if run_captured(self, kind, LR, HR, Z):
                    return
            self._train_step_body(kind, LR, HR, Z)
            return
        self.make_new_labels()
        fake_HR = self.update_G(LR, HR, Z, it_host, False)
        self.update_D(HR, fake_HR, it_host, False)
        if HR.is_cuda and HR.shape[1] == 3:
            # PSNR of SR and of the trilinear baseline + the trilinear pixel loss from ONE fused pass
            # (wind_field_GAN_3D.py:597-618, 730-770); everything stays on the device
            sums = ops.validation_metrics(HR, fake_HR, LR)
            vox = float(HR.shape[0] * HR.shape[2] * HR.shape[3] * HR.shape[4])
            mse = (sums[:2] / vox).to(torch.float32)
            psnr = 10.0 * torch.log10(self.max_diff_squared / (mse + self.epsilon_PSNR))
            self.metrics_dict["val_PSNR"], self.metrics_dict["Trilinear_PSNR"] = psnr[0], psnr[1]
            self.metrics_dict["trilinear_pix_loss"] = ((sums[2] if self.pixel_criterion != "l2" else sums[1])
                                                       / (3.0 * vox)).to(torch.float32)
            return
        (self.metrics_dict["val_PSNR"], self.metrics_dict["Trilinear_PSNR"]) = compute_PSNR_for_SR_and_trilinear(
            LR, HR, fake_HR, self.max_diff_squared, self.epsilon_PSNR, interpolate=True, device=self.device,
            scale=self.cfg.scale)
        tri = F.interpolate(LR[:, :3], scale_factor=(self.cfg.scale, self.cfg.scale, 1), mode="trilinear",
                            align_corners=True)
        self.metrics_dict["trilinear_pix_loss"] = (F.l1_loss(HR, tri) if self.pixel_criterion != "l2"
                                                   else F.mse_loss(HR, tri))

    def optimize_parameters(self, LR, HR, Z, it):
        self.compute_losses_and_optimize(LR, HR, Z, it, training_iteration=True)
        # The step has no host synchronisation, so the host could run arbitrarily far ahead of the device.  Under
        # data parallelism that let the ranks drift apart until the NCCL kernels of one rank spun on the other's
        # backlog (measured: 100-170 ms stalls in ~1 of 6 steps at 2 GPUs).  Bound the run-ahead to two steps: wait
        # for the step before the previous one — it has normally finished already, so this costs nothing.
        if HR.is_cuda:
            ev = torch.cuda.Event()
            ev.record()
            self._inflight.append(ev)
            if len(self._inflight) > 2:
                self._inflight.pop(0).synchronize()

    def validation(self, LR, HR, Z, it):
        self.compute_losses_and_optimize(LR, HR, Z, it, training_iteration=False)

    # ---------------------------------------------------------------------------------------------------
    def make_new_labels(self, it=None):
        """Real / fake target vectors (wind_field_GAN_3D.py:627-678): optional flip, one-sided smoothing that
        anneals 0.9 -> 1.0 over ``niter``, optional Gaussian label noise — built from the device scalars written by
        ``_write_scalars`` (``it`` given: write them first; kept for callers of the reference's signature)."""
        if it is not None:
            self._write_scalars(int(it))
        t = self.cfg.training
        real, fake = self._scalars[self._S_REAL:self._S_REAL + 1], self._scalars[self._S_FAKE:self._S_FAKE + 1]
        if t.flip_labels:
            real, fake = fake, real
        if t.use_noisy_labels:
            mk = lambda base: torch.clamp(torch.randn(self.batch_size, device=self.device) * 0.05 + base, 0.0, 1.0)
            self.HR_labels, self.fake_HR_labels = mk(real).squeeze(), mk(fake).squeeze()
            self._labels_point_nine = torch.all(self.HR_labels == 0.9).to(torch.float32)
        else:
            self.HR_labels = real.expand(self.batch_size).squeeze()
            self.fake_HR_labels = fake.expand(self.batch_size).squeeze()
            self._labels_point_nine = self._scalars[self._S_POINT_NINE]

    # ---------------------------------------------------------------------------------------------------
    def get_G_train_loss_dict_ref(self):
        return self.train_G_loss_dict

    def get_G_val_loss_dict_ref(self):
        return self.validation_G_loss_dict

    def get_D_loss_dict_ref(self):
        return self.D_loss_dict

    def get_hist_dict_ref(self):
        return self.hist_dict

    def get_metrics_dict_ref(self):
        return self.metrics_dict

    def update_learning_rate(self):
        for s in self.schedulers:
            s.step()

    def count_params(self):
        return (sum(p.numel() for p in self.G.parameters()), sum(p.numel() for p in self.D.parameters()))

    def count_trainable_params(self):
        return (sum(p.numel() for p in self.G.parameters() if p.requires_grad),
                sum(p.numel() for p in self.D.parameters() if p.requires_grad))

    def __str__(self):
        g, d = self.count_params()
        gt, dt = self.count_trainable_params()
        bar = "*---------------*"
        return (f"{bar}\nGenerator:\n{g} params, {gt} trainable\n\n{self.G}\n\n"
                f"{bar}\nDiscriminator:\n{d} params, {dt} trainable\n\n{self.D}\n")


def calculate_PSNR(HR, fake_HR, max_diff_squared=torch.tensor(4.0), epsilon_PSNR=torch.tensor(1e-8),
                   device=torch.device("cpu")):
    """10 log10(max^2 / (batch-mean squared error + eps)) (wind_field_GAN_3D.py:730-742), kept on device."""
    voxels = HR.shape[0] * HR.shape[2] * HR.shape[3] * HR.shape[4]
    mse = torch.sum((HR - fake_HR) ** 2) / voxels
    return 10.0 * torch.log10(max_diff_squared / (mse + epsilon_PSNR))


def compute_PSNR_for_SR_and_trilinear(LR, HR, fake_HR, max_diff_squared, epsilon_PSNR, interpolate=False,
                                      device=torch.device("cpu"), scale=4):
    psnr = calculate_PSNR(HR, fake_HR, max_diff_squared, epsilon_PSNR, device=device)
    if not interpolate:
        return psnr
    tri = F.interpolate(LR[:, :3], scale_factor=(scale, scale, 1), mode="trilinear", align_corners=True)
    return psnr, calculate_PSNR(HR, tri, max_diff_squared, epsilon_PSNR, device=device)


def get_norm_factors_of_gradients(HR_wind_gradient, SR_wind_gradient):
    """The four loss normalisers from materialised Jacobians (wind_field_GAN_3D.py:773-814); the training step
    itself uses the fused slots instead.  Note the z-gradient max is taken WITHOUT abs, like the reference."""

    def div(g):
        return g[:, 0] + g[:, 4] + g[:, 8]

    def dxy(g):
        return g[:, 0] + g[:, 4]

    pairs = [
        (HR_wind_gradient[:, :6].abs().max(), SR_wind_gradient[:, :6].abs().max()),
        (HR_wind_gradient[:, 6:].max(), SR_wind_gradient[:, 6:].max()),
        (div(HR_wind_gradient).abs().max(), div(SR_wind_gradient).abs().max()),
        (dxy(HR_wind_gradient).abs().max(), dxy(SR_wind_gradient).abs().max()),
    ]
    return [torch.max(h, s / 100) for h, s in pairs]
