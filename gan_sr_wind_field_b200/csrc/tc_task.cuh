// tc_task.cuh — optional override of the conv geometry seen by the tcgen05 conv kernels: lets one launch compute a
// sub-problem (a parity class of a strided data-gradient) with its own tap set and a strided destination grid.
#pragma once
namespace ws {
struct TcOverride {
  int DX, DY, DZ;        // destination sub-grid extents
  int SX, SY, SZ;        // source tensor extents
  int kx, ky, kz;        // taps of this sub-problem
  int px, py, pz;        // src coordinate of (dst d, tap t) = d - p + t
  int ck, cn;            // reduction / destination channels
  int tap_base, taps_total;  // slice [tap_base, tap_base + kx*ky*kz) of the packed weight's tap dimension
  int omx, oax, omy, oay, omz, oaz;  // destination voxel = (d*om + oa) per axis ...
  int ODY, ODZ;          // ... inside a grid with these y / z extents
};
}  // namespace ws
