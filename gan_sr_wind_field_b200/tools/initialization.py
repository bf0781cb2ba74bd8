"""Weight initialisation with the reference's semantics (tools/initialization.py:15-34): Kaiming-normal
(fan-in, a=0) times ``scale`` for every module whose *class name* is Conv2d/Conv3d or contains "Linear",
biases zeroed; BatchNorm is left untouched (the reference's "BatchNorm3D" spelling never matches)."""
import functools

from torch.nn import init


def init_kaiming(m, scale=1):
    name = type(m).__name__
    if name in ("Conv2d", "Conv3d") or "Linear" in name:
        init.kaiming_normal_(m.weight.data, a=0, mode="fan_in")
        m.weight.data *= scale
        if m.bias is not None:
            m.bias.data.zero_()
    elif name in ("BatchNorm2d", "BatchNorm3D"):
        init.constant_(m.weight.data, 1.0)
        init.constant_(m.bias.data, 0.0)


def init_weights(m, scale=1):
    m.apply(functools.partial(init_kaiming, scale=scale))
