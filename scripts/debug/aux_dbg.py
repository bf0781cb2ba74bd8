"""Debug: allocator behaviour of the RRDB backward with the auxiliary weight-gradient stream."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from gan_sr_wind_field_b200 import ops
from gan_sr_wind_field_b200.CNN_models.torch_blocks import RRDB
ops.set_precision("bf16")
torch.manual_seed(0)
blk = RRDB(128, 32, 5, 1, lrelu_negative_slope=0.2, RDB_residual_scaling=0.2, RRDB_residual_scaling=0.2, mode="3D").cuda()
x = torch.randn(8, 128, 16, 16, 10, device="cuda").contiguous(memory_format=torch.channels_last_3d).requires_grad_(True)
for it in range(8):
    st0 = torch.cuda.memory_stats()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    y = blk(x)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    y.backward(torch.ones_like(y))
    t2 = time.perf_counter()
    torch.cuda.synchronize(); t3 = time.perf_counter()
    st1 = torch.cuda.memory_stats()
    print(f"it {it}: fwd {1e3*(t1-t0):.2f} ms, bwd host {1e3*(t2-t1):.2f} ms, bwd total {1e3*(t3-t1):.2f} ms, "
          f"cudaMalloc {st1['num_device_alloc']-st0['num_device_alloc']}, cudaFree {st1['num_device_free']-st0['num_device_free']}, "
          f"reserved {st1['reserved_bytes.all.current']/1e6:.0f} MB")
