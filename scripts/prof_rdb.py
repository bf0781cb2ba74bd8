"""Time one trunk RRDB (3 residual dense blocks, 128 features, gc 32, 16x16x10, B=8, bf16) forward and backward,
optionally sweeping the wgrad split (WS_WGRAD_FORCE) in-process."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from gan_sr_wind_field_b200 import ops
from gan_sr_wind_field_b200.CNN_models.torch_blocks import RRDB

ops.set_precision("bf16")
torch.manual_seed(0)
blk = RRDB(128, 32, 5, 1, lrelu_negative_slope=0.2, RDB_residual_scaling=0.2, RRDB_residual_scaling=0.2,
           mode="3D").cuda()
x = torch.randn(8, 128, 16, 16, 10, device="cuda").contiguous(memory_format=torch.channels_last_3d).requires_grad_(True)


def run(reps=20):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tf = tb = 0.0
    for i in range(reps + 3):
        ev[0].record()
        y = blk(x)
        ev[1].record()
        y.backward(torch.ones_like(y))
        ev[2].record()
        torch.cuda.synchronize()
        if i >= 3:
            tf += ev[0].elapsed_time(ev[1]); tb += ev[1].elapsed_time(ev[2])
    return tf / reps * 1e3 / 3, tb / reps * 1e3 / 3


f, b = run()
print(f"default: fwd {f:.0f} us/RDB, bwd {b:.0f} us/RDB")
for cfg in sys.argv[1:]:
    os.environ["WS_WGRAD_FORCE"] = cfg
    f, b = run()
    print(f"WS_WGRAD_FORCE={cfg}: fwd {f:.0f} us/RDB, bwd {b:.0f} us/RDB")
