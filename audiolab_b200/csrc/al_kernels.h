// Internal kernel parameter blocks and launchers shared by the .cu files (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "al_fft.cuh"

#define AL_DYN_SMEM(T, name) extern __shared__ __align__(16) unsigned char name##_raw_[]; T* name = reinterpret_cast<T*>(name##_raw_)

namespace al {

void count_launch();

// "Has this been done on the CURRENT device?"  Function attributes (the > 48 KB shared memory opt-in) and the SM
// count are per device; a process that drives several GPUs must not reuse the first device's answers.
struct PerDeviceOnce {
    unsigned long long done = 0;
    bool needed() const {
        int d = 0;
        cudaGetDevice(&d);
        return ((__atomic_load_n(&done, __ATOMIC_ACQUIRE) >> (d & 63)) & 1ull) == 0;
    }
    void mark() {
        int d = 0;
        cudaGetDevice(&d);
        __atomic_fetch_or(&done, 1ull << (d & 63), __ATOMIC_RELEASE);
    }
};
int sm_count();   // of the current device (cached per device)

// [emul-begin]

struct StftParams {
    const float* track;
    long long n_valid;
    long long ch_stride;
    int channels;
    const long long* chunk_offsets;
    long long off0, off_step;
    int chunk_len;
    int center;
    int hop;
    int n_frames;
    const float* window;   // [N] analysis window (already scaled)
    const float2* tw;      // [32*32]  exp(-2 pi i k1 n2 / 1024)
    const float2* ctw;     // [(D-1)*513] exp(-2 pi i r kappa / N), r = 1..D-1
    float* spec;
    int layout;
    int n_bins_out;
    int zero_low_bins;
    // filled by the launcher
    int ps;
    int rounds_per_cta;
    int tiles;
};

struct IstftParams {
    const float* spec;
    const float* mask;
    int layout;
    int n_bins_in;
    int n_frames_in;
    int frame_pad;
    int n_frames_total;    // n_frames_in + 2*frame_pad
    int stems, channels;
    int spec_has_stems;
    int zero_low_bins;
    int hop;
    const float* window;   // [N] synthesis window * 1/N (* sqrt(N) if normalized)
    const float2* tw;
    const float2* ctw;     // forward combine twiddles; the inverse uses their conjugate
    const float* inv_env;  // [(n_frames_total-1)*hop + N]
    int out_start;
    int out_len;
    const float* weight;
    float* dst;
    long long dst_ch_stride, dst_chunk_stride;
    const long long* dst_offsets;
    long long dst_off0, dst_off_step;
    long long dst_limit;
    // filled by the launcher
    int hops_per_cta;
    int segs;
};

// stereo / n_fft 2048 / bin-innermost fast path (al_stft_pk.cu)
struct StftPkParams {
    const float* track;
    long long n_valid;
    long long ch_stride;
    const long long* chunk_offsets;
    long long off0, off_step;
    int n_chunks;
    int chunk_len;
    int center;
    int hop;
    int n_frames;
    const float* window;      // [2048] analysis window (already scaled)
    const float2* tw;         // [32*32]
    const float2* ctw_half;   // [544] 0.5 * exp(-2 pi i k / 2048), zero padded beyond 512
    float* spec;
    int layout;               // 0 or 3
    int n_bins_out;
    int zero_low_bins;
    int aligned;              // track and ch_stride allow 128-bit loads
    // filled by the launcher
    int sp;
    int n_stages;
    int tiles_per_chunk;
    int total_tiles;
    long long* prof;          // AL_PK_PROF builds only: per-warp phase cycle counters
};

// stereo / n_fft 2048 / FRAME_INTERLEAVED fast path of K2 (al_istft_pk.cu)
struct IstftPkParams {
    const float4* spec;       // [chunks (* stems)][T][1025] (L.re, L.im, R.re, R.im)
    const float4* mask;       // [chunks * stems][T][1025] or NULL
    int n_frames;             // T (no frame padding on this path)
    int stems;
    int spec_has_stems;
    int hop;
    const float* window;      // [2048] synthesis window * 1/N
    const float2* tw;         // [32*32]
    const float2* ctw;        // [1024] exp(-2 pi i k / 2048)
    const float* inv_env;     // [(T-1)*hop + 2048]
    int out_start;
    int out_len;
    const float* weight;
    float* dst;
    long long dst_ch_stride, dst_chunk_stride;
    const long long* dst_offsets;
    long long dst_off0, dst_off_step;
    long long dst_limit;
    int ola_fast;             // interior rounds take the predicate-free overlap-add (AL_IP_OLAFAST=0 disables)
    int l2_prefetch;          // bulk L2 prefetch of the next round's rows (AL_IP_L2PF=0 disables)
    int tok_hint_ns;          // istft_pk5_kernel: suspend-time hint of the token waits (0 = plain try_wait; AL_IP_HINT)
    int tok_relaxed;          // istft_pk5_kernel: token arrivals with .relaxed semantics (experiment; AL_IP_RELAXED)
    // filled by the launcher
    int hops_per_cta;
    int segs;
};

// [emul-end]

cudaError_t launch_stft(const StftParams& p, int n_fft, int rows, cudaStream_t stream);
cudaError_t launch_istft_pk(const IstftPkParams& p, int n_chunks, cudaStream_t stream);
cudaError_t launch_stft_pk(const StftPkParams& p, cudaStream_t stream);
cudaError_t launch_istft(const IstftParams& p, int n_fft, int n_chunks, cudaStream_t stream);

cudaError_t launch_ola_gather(const float* chunks, int n_chunks, int data_chunk0, int rows, int chunk_len,
                              const long long* offsets, const int* mult, const float* wtab,
                              const int* tab_id, long long n_total, long long p0, long long p1,
                              const float* halo_in, int raw_out, float eps, float scale, float* track,
                              long long track_stride, cudaStream_t stream);

cudaError_t launch_resample(const float* in, long long in_stride, float* out, long long out_stride,
                            int rows, long long n_in, long long n_out, int up, int down,
                            const float* taps, int n_taps, cudaStream_t stream);

cudaError_t launch_sub(const float* a, const float* b, float* out, long long n, cudaStream_t stream);

cudaError_t launch_rmsnorm_bf16(void* x, const float* gamma, const float* bias, void* out, long long n_rows, int dim,
                                float scale, float eps, cudaStream_t stream);
cudaError_t launch_rotary_bf16(void* q, void* k, const float* cs, long long n_rows, int heads, int dim_head,
                               long long pos_div, int pos_mod, cudaStream_t stream);
cudaError_t launch_gate_bf16(void* o, const void* gates, long long n_rows, int heads, int dim_head, int gate_ld, int fp16,
                             cudaStream_t stream);

cudaError_t launch_gelu_bf16(void* x, long long n, cudaStream_t stream);
cudaError_t launch_band_attn_bf16(const void* q, const void* k, const void* v, void* o, const void* gates, const float* cos_sin,
                                  long long n_seq, int F, int heads, float scale, int gate_ld, int fp16, cudaStream_t stream);

// csrc/al_fattn.cu: returns NULL on success, else a static message (and the CUDA error, if that is what failed, in *cuda_err)
const char* launch_time_attention(const void* q, const void* k, const void* v, void* o, const void* gates, long long gate_ld,
                                  long long n_batch, int seq_len, int inner, int heads, int dim_head, float scale, int fp16,
                                  cudaStream_t stream, cudaError_t* cuda_err);

cudaError_t launch_env(const float* window_raw, int n_fft, int hop, int n_frames_total, float* inv_env,
                       cudaStream_t stream);

}  // namespace al
