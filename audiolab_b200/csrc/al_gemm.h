// Internal argument blocks of the tcgen05 GEMM (al_gemm.cu); the public call is al_gemm_bf16 (include/audiolab_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/audiolab_b200.h"

namespace al {

enum { EPI_BF16 = AL_GEMM_EPI_BF16, EPI_RES = AL_GEMM_EPI_RESIDUAL, EPI_GLU = AL_GEMM_EPI_GLU };
enum { ACT_NONE = AL_GEMM_ACT_NONE, ACT_GELU = AL_GEMM_ACT_GELU, ACT_TANH = AL_GEMM_ACT_TANH };

typedef al_gemm_args GemmCall;

// What the kernel needs besides the tensor maps.
struct GemmArgs {
    int M, N, K, groups;
    int m_tiles, n_tiles;
    const float* bias;
    const float* row_ss;
    int ss_parts;
    float ss_scale, ss_eps;
    const float* cos_sin;
    long long pos_div;
    int pos_mod, rot_cols;
    int act;
    int out_split;
    float* ss_out;
    int max_ctas;
    long long side_rs, side_gs;   // per-row side arrays (row_ss, ss_out): row index = group * side_gs + row * side_rs
    int accumulate;               // EPI_RES: 0 = start the stream (no residual read)
    int fp16;                     // 16-bit operands / outputs are IEEE half instead of bfloat16
    int pf_x;                     // EPI_RES: prefetch the fp32 residual tile into L2 when its main loop starts
};

// Returns NULL on success, else a static message (and the CUDA error, if that is what failed, in *cuda_err).
const char* launch_gemm_bf16(const GemmCall& c, cudaStream_t stream, cudaError_t* cuda_err);

cudaError_t launch_band_norm(const float* x, long long ldx, const float* gamma, const int* band_off, int n_bands, void* out,
                             long long ldo, long long n_rows, float eps, int fp16, cudaStream_t stream);

cudaError_t launch_resid_prepare(const float* x_in, const float* bias, const float* gamma, float* x32, void* xb, float* ss,
                                 long long n_rows, int dim, int ss_parts, float eps, int fp16, cudaStream_t stream);

}  // namespace al
