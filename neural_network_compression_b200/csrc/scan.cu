// scan.cu -- multi-CTA exclusive prefix sums of the small per-tile arrays (survivor counts per reduction tile,
// fixed-point sums per 1024-key tile of the sorted survivors).  A single-CTA scan of 3e5 entries took 0.3 ms.
#include <algorithm>

#include "common.cuh"
#include "internal.h"
#include "table.cuh"

namespace nnc {

template <class TIn, class TOut>
static void exclusive_scan(nnc_ctx *ctx, const TIn *d_in, long long n, TOut *d_out) {
    if (n <= 0) {
        NNC_CUDA(cudaMemsetAsync(d_out, 0, sizeof(TOut), ctx->stream));
        return;
    }
    int chunks = (int)std::min<long long>(std::min(1024, ctx->sm_count * 2), (n + 4095) / 4096);
    long long chunk = (n + chunks - 1) / chunks;
    chunk = (chunk + 4095) / 4096 * 4096;
    chunks = (int)((n + chunk - 1) / chunk);
    TOut *chunk_sum = arena_alloc_t<TOut>(ctx, (size_t)chunks + 1);
    NNC_LAUNCH(ctx, (scan_chunk_sum_kernel<TIn, TOut>), chunks, 1024, 0, d_in, n, chunk, chunk_sum);
    NNC_LAUNCH(ctx, (scan_chunk_apply_kernel<TIn, TOut>), chunks, 1024, 0, d_in, n, chunk, chunk_sum, d_out);
}

void exclusive_scan_i64(nnc_ctx *ctx, const long long *d_in, long long n, long long *d_out) {
    exclusive_scan<long long, long long>(ctx, d_in, n, d_out);
}
void exclusive_scan_u32_u64(nnc_ctx *ctx, const unsigned int *d_in, long long n, unsigned long long *d_out) {
    exclusive_scan<unsigned int, unsigned long long>(ctx, d_in, n, d_out);
}

}  // namespace nnc
