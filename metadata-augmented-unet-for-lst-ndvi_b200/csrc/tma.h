// tma.h -- host-side CUtensorMap construction (driver entry point fetched at run time, so the
// library links against libcudart only and loads on a machine without libcuda).
#pragma once
#include <cuda.h>
#include "common.h"

namespace mau {

// rank-N tiled tensor map over a bf16/fp32 tensor.  dims[0] is the contiguous dimension.
// strides_bytes[i] is the stride of dims[i+1] (rank-1 entries).  128B swizzle when swizzle128.
int make_tensor_map(CUtensorMap* out, int dtype, int rank, void* base, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, bool swizzle128);

// NHWC activation view {C, W, H, B} with box {bc, bw, bh, 1}
int make_nhwc_map(CUtensorMap* out, int dtype, const View& v, int box_c, int box_w, int box_h);

}  // namespace mau
