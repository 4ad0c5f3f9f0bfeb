// Host-side helpers shared by the C-ABI entry points: error reporting and TMA tensor-map encoding.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/backpack_b200.h"

namespace bp {

// printf-style; stores the message returned by bp_last_error() (thread local) and returns `code`.
int fail(int code, const char* fmt, ...);

// Encode a tiled tensor map without linking libcuda: the driver entry point is resolved through
// cudaGetDriverEntryPoint the first time it is needed.  `dims`/`box` are innermost-first,
// `strides_bytes` has rank-1 entries (stride of dims 1..rank-1).  Returns 0 or a negative bp error.
int encode_tensor_map(CUtensorMap* out, int elem_bytes_log2_dtype /*bp_dtype_t*/, int rank, const void* base,
                      const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                      bool swizzle128);

int check_launch(const char* what);

// attention dropout helpers (bp_fmha_bwd.cu): column-word table launch, 8-bit threshold of a probability
int launch_drop_col_table(uint32_t* table, uint64_t seed, int32_t bh, int32_t s_pad_k, cudaStream_t st);
int drop_threshold(float p);
// residual dropout inside the LayerNorm kernels (bp_layer_norm.cu): validates p, derives the hash base of `seed`, the
// 8-bit threshold shifted to the top byte (0: no dropout) and the 1 / (1 - p_effective) scale
int ln_drop_args(float p, uint64_t seed, int64_t rows, const char* fn, uint32_t* base, uint32_t* thr24, float* scale);

inline int dtype_size(int dtype) { return dtype == BP_DTYPE_F32 ? 4 : 2; }

}  // namespace bp
