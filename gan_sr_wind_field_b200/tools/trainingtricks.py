"""GAN training tricks with the reference's semantics (tools/trainingtricks.py:18-58), kept on the device
(the reference samples label noise on the CPU and copies it over: two H2D copies per step)."""
import torch


def noisy_labels(label_type, batch_size, noise_stddev=0.05, false_label_val=0.0, true_label_val=1.0,
                 val_lower_lim=0.0, val_upper_lim=1.0, device=torch.device("cpu")):
    """Gaussian-perturbed real/fake label vector of length ``batch_size``, clamped to [lower, upper]."""
    std = float(noise_stddev)
    base = float(true_label_val if label_type else false_label_val)
    lo, hi = float(val_lower_lim), float(val_upper_lim)
    if std > 0.0:
        label = torch.randn(int(batch_size), device=device) * std + base
        return torch.clamp(label, min=lo, max=hi)
    # no noise: a constant vector, filled on the device (no host RNG round trip, no H2D copy)
    return torch.full((int(batch_size),), min(max(base, lo), hi), device=device)


def instance_noise(sigma_base, shape, it, niter, device=torch.device("cpu")):
    """U[0,1) * sqrt(sigma_base * (1 - (it-1)/niter)) — uniform, as the reference actually draws it
    (trainingtricks.py:56)."""
    noise = torch.rand(shape, device=device)
    return noise * torch.sqrt(sigma_base * (1 - (it - 1) / niter))
