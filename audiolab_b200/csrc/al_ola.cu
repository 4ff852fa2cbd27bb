// K2b: deterministic windowed overlap-add of chunk outputs into a track (gather form), and the
// elementwise complement stem.  Replaces the `result += x*window; counter += window; result/counter`
// loops of upstream MDXSeparator.demix / MDXCSeparator.demix / demucs.apply.
//
// HBM-bound: reads every chunk sample once (4 B) and writes every track sample once.
// One thread owns one (row, position); it adds the covering chunks in ascending chunk order, so the
// sum is bit-identical however the chunks were batched or sharded.  The weight sum ("counter") is
// recomputed from the same tables instead of being stored.
#include "al_kernels.h"

namespace al {

__global__ void __launch_bounds__(256)
ola_gather_kernel(const float* __restrict__ chunks, int n_chunks, int data_chunk0, int rows, int chunk_len,
                  const long long* __restrict__ offsets, const int* __restrict__ mult,
                  const float* __restrict__ wtab, const int* __restrict__ tab_id, long long n_total,
                  long long p0, long long p1, const float* __restrict__ halo_in, int raw_out, float eps,
                  float scale, float* __restrict__ track, long long track_stride) {
    const long long p = p0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= p1) return;
    // first chunk whose end is beyond p: offsets ascending, so binary search on off_c + chunk_len > p
    int lo = 0, hi = n_chunks;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(offsets + mid) + chunk_len > p) hi = mid; else lo = mid + 1;
    }
    const int c_first = lo;
    for (int r = blockIdx.y; r < rows; r += gridDim.y) {
        float acc = halo_in ? __ldg(halo_in + (long long)r * (p1 - p0) + (p - p0)) : 0.f;
        float wsum = 0.f;
        for (int c = c_first; c < n_chunks; ++c) {
            const long long off = __ldg(offsets + c);
            if (off > p) break;
            const long long j = p - off;
            const long long len = min((long long)chunk_len, n_total - off);
            if (j >= len) continue;
            float w = 1.f;
            if (wtab) w = __ldg(wtab + (long long)(tab_id ? __ldg(tab_id + c) : 0) * chunk_len + j);
            const int m = mult ? __ldg(mult + c) : 1;
            // chunks before data_chunk0 belong to the left neighbour: their partial sums arrive through
            // halo_in, only their weights are counted here
            const float x = c >= data_chunk0 ? __ldg(chunks + ((long long)(c - data_chunk0) * rows + r) * chunk_len + j) : 0.f;
            for (int k = 0; k < m; ++k) {   // the reference re-adds a tail chunk m times; keep its rounding
                acc += x * w;
                wsum += w;
            }
        }
        float v;
        if (raw_out) v = acc;
        else v = scale * acc / fmaxf(wsum, eps);
        track[(long long)r * track_stride + p] = v;
    }
}

cudaError_t launch_ola_gather(const float* chunks, int n_chunks, int data_chunk0, int rows, int chunk_len,
                              const long long* offsets, const int* mult, const float* wtab,
                              const int* tab_id, long long n_total, long long p0, long long p1,
                              const float* halo_in, int raw_out, float eps, float scale, float* track,
                              long long track_stride, cudaStream_t stream) {
    if (p1 <= p0) return cudaSuccess;
    const long long n = p1 - p0;
    dim3 grid((unsigned)((n + 255) / 256), (unsigned)min(rows, 8));
    ola_gather_kernel<<<grid, 256, 0, stream>>>(chunks, n_chunks, data_chunk0, rows, chunk_len, offsets, mult, wtab, tab_id,
                                                n_total, p0, p1, halo_in, raw_out, eps, scale, track,
                                                track_stride);
    count_launch();
    return cudaGetLastError();
}

__global__ void __launch_bounds__(256)
sub_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, long long n) {
    const long long i4 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i4 + 3 < n) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(a + i4));
        const float4 y = __ldg(reinterpret_cast<const float4*>(b + i4));
        *reinterpret_cast<float4*>(out + i4) = make_float4(x.x - y.x, x.y - y.y, x.z - y.z, x.w - y.w);
    } else {
        for (long long i = i4; i < n; ++i) out[i] = a[i] - b[i];
    }
}

cudaError_t launch_sub(const float* a, const float* b, float* out, long long n, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    if ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(out)) & 15)
        return cudaErrorMisalignedAddress;
    sub_kernel<<<(unsigned)((n + 1023) / 1024), 256, 0, stream>>>(a, b, out, n);
    count_launch();
    return cudaGetLastError();
}

}  // namespace al
