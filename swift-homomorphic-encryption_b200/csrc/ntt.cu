// ntt.cu -- dispatch between the NTT implementations.
//   HECUDA_NTT_IMPL=simple forces the generic radix-2 shared-memory kernel (ntt_simple.cu);
//   otherwise the register-tiled kernel (ntt_fast.cu) is used for the sizes it supports.
#include <cstdlib>
#include <cstring>

#include "kernels.cuh"

namespace hecuda {

cudaError_t launch_ntt_forward(const Context &ctx, const NttRowMap &map, const u64 *in, u64 *out, int64_t rows,
                               cudaStream_t stream) {
    return launch_ntt_forward_simple(ctx, map, in, out, rows, stream);
}

cudaError_t launch_ntt_inverse(const Context &ctx, const NttRowMap &map, const u64 *in, u64 *out, int64_t rows,
                               bool scale_t, cudaStream_t stream) {
    return launch_ntt_inverse_simple(ctx, map, in, out, rows, scale_t, stream);
}

}  // namespace hecuda
