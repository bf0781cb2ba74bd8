// tmap.cuh — cached cuTensorMapEncodeTiled descriptors shared by the tcgen05 kernels.
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace ws {

struct MapKey {
  uintptr_t ptr;
  uint64_t dims[5];
  uint64_t strides[4];  // bytes, dims 1..rank-1
  uint32_t box[5];
  uint32_t estr[5];
  uint32_t rank;
  uint32_t dtype;
};
static_assert(sizeof(MapKey) % 8 == 0, "MapKey must hash as 64-bit words");
constexpr uint32_t kMapSwizzle64 = 1u << 8;  // OR into MapKey::dtype: SWIZZLE_64B (64-byte rows) instead of 128B
// OR into MapKey::dtype: SWIZZLE_128B with 32-byte atoms — the only shared-memory layout tcgen05 accepts for MN-major
// (transposed) TF32 operands (UMMA layout type 1, "128B_BASE32B": the 4 x 32-byte chunks of a 128-byte row are XORed
// with the row index mod 4, pattern period 512 B)
constexpr uint32_t kMapSwizzle128Atom32 = 1u << 9;

// SWIZZLE_128B (or, with kMapSwizzle64, SWIZZLE_64B) tiled tensor map for `key` (memset the key to 0 before filling it). Returns 0 on success.
int get_tensor_map(const MapKey& key, CUtensorMap* out);

}  // namespace ws
