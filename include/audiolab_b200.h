/*
 * audiolab_b200.h -- C ABI of libaudiolab_b200.so (hand-written sm_100a kernels for the
 * AudioLab source-separation spectral hot path).
 *
 * The reference (d8ahazard/AudioLab) has NO FFI: its hot path is Python calling
 * torch.stft / torch.istft / numpy inside the third-party `audio_separator` package.
 * Each entry point below therefore cites the reference *Python* interface it replaces
 * (paths relative to /root/reference); INTEGRATION.md shows the ctypes binding and the
 * monkey-patch a maintainer would add (same mechanism as handlers/patch_separate.py:71-78).
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error (AL_E_*); al_last_error() gives the
 *     message for the calling thread.  Nothing throws.  Nothing allocates caller-visible memory.
 *   - all data pointers are DEVICE pointers owned by the caller (fp32 unless stated); work is
 *     enqueued on `stream` (a cudaStream_t passed as void*) and is asynchronous.
 *   - plans own small read-only device tables (twiddles, windows, OLA envelopes); they are
 *     immutable after creation except for an internal mutex-guarded envelope cache, so one plan
 *     may be used from several streams / threads.
 *   - spectrogram layouts (`layout`), `Fo` = number of stored bins (dim_f crop, <= n_fft/2+1):
 *       AL_LAYOUT_FRAME_MAJOR 0   complex64 [rows, T, Fo]          (rows = chunk*channels + ch)
 *       AL_LAYOUT_BIN_MAJOR   1   complex64 [rows, Fo, T]          == torch.stft(return_complex=True)
 *       AL_LAYOUT_CAC         2   float32   [chunks, channels*2, Fo, T]   "complex as channels",
 *                                 channel order (L.re, L.im, R.re, R.im)  == mdxnet.py:51-56
 *       AL_LAYOUT_FRAME_INTERLEAVED 3 complex64 [groups, T, Fo, channels]  (group = chunk, or
 *                                 chunk*stems + stem) == upstream BSRoformer 'b t (f s c)': the band-split
 *                                 input and the mask-estimator output with no permute copy
 */
#ifndef AUDIOLAB_B200_H_
#define AUDIOLAB_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AL_OK 0
#define AL_E_ARG (-1)     /* bad argument */
#define AL_E_CUDA (-2)    /* CUDA runtime error */
#define AL_E_UNSUPPORTED (-3)

#define AL_LAYOUT_FRAME_MAJOR 0
#define AL_LAYOUT_BIN_MAJOR 1
#define AL_LAYOUT_CAC 2
#define AL_LAYOUT_FRAME_INTERLEAVED 3

typedef struct al_plan al_plan;

/* Library version, e.g. 100 = 0.1.0. */
int al_version(void);

/* Message of the last error raised on the calling thread ("" if none). */
const char* al_last_error(void);

/* Number of kernels this library has launched in this process (bench.py's gpu_launches). */
int64_t al_launch_count(void);

/*
 * Plan for one (n_fft, hop, window, normalized) STFT configuration.
 *   n_fft in {2048, 4096, 6144}; window = HOST pointer to n_fft floats, or NULL for the
 *   periodic Hann window torch.hann_window(n_fft) every reference call site uses
 *   (modules/rvc/infer/modules/uvr5/mdxnet.py:27); normalized != 0 reproduces
 *   torch.stft(normalized=True) (HTDemucs `spectro`).
 */
int al_plan_create(int n_fft, int hop, const float* window_host, int normalized, al_plan** out);
int al_plan_destroy(al_plan* plan);

/*
 * K1  al_stft -- fused pad-and-chunk + reflect-pad + framing + window + R2C FFT + layout.
 * Replaces: torch.stft(x, n_fft, hop, window, center=True) plus the layout shuffles around it:
 *   modules/rvc/infer/modules/uvr5/mdxnet.py:41-56 (ConvTDFNetTrim.stft), the chunk gather of
 *   mdxnet.py:152-164 (Predictor.demix_base), and upstream audio_separator STFT.__call__ /
 *   BSRoformer.forward stft / HTDemucs._spec.
 *
 * track         [channels, >= n_valid] fp32, channel stride `ch_stride` floats; samples outside
 *               [0, n_valid) read as zero (the reference's explicit zero padding).
 * chunk c       covers track[off_c : off_c + chunk_len], off_c = chunk_offsets[c] (device int64
 *               array) if non-NULL else off0 + c*off_step; off_c may be negative.
 * center_pad    samples of reflect padding before chunk sample 0 (n_fft/2 for center=True;
 *               HTDemucs._spec folded with its frame crop gives 3*hop/2). Requires chunk_len > center_pad.
 * n_frames      frames to emit: frame t covers chunk samples [t*hop - center_pad, ... + n_fft).
 * spec          output, `layout` above with T = n_frames, Fo = n_bins_out; bins < zero_low_bins are
 *               written as 0 (upstream MDXSeparator zeroes spek[:, :, :3, :]).
 */
int al_stft(const al_plan* plan, const float* track, int64_t n_valid, int64_t ch_stride, int channels,
            const int64_t* chunk_offsets, int64_t off0, int64_t off_step, int n_chunks, int chunk_len,
            int center_pad, int n_frames, float* spec, int layout, int n_bins_out, int zero_low_bins,
            void* stream);

/*
 * K2  al_istft -- fused (complex mask (.) spec | CaC -> complex) + zero freq-pad + C2R iFFT +
 *                 synthesis window + overlap-add over frames + / sum(window^2) + centre trim
 *                 [+ per-sample chunk weight] [+ trim-and-concat placement into a track buffer].
 * Replaces: torch.istft(...) and its surroundings: mdxnet.py:58-75 (ConvTDFNetTrim.istft),
 *   mdxnet.py:178-183 (trim + concat), BSRoformer.forward `stft_repr * mask` + istft,
 *   HTDemucs._mask/_ispec.
 *
 * spec          `layout`, T = n_frames_in, Fo = n_bins_in, rows = chunk*channels + ch; bins >= n_bins_in
 *               are zero (mdxnet.py:34-36,59-64 freq_pad).  When spec_has_stems != 0 the spectrogram
 *               carries the stem axis itself: rows = (chunk*stems + stem)*channels + ch (HTDemucs).
 * mask          NULL, or complex64 in the same complex `layout` (0, 1 or 3) with rows
 *               (chunk*stems + stem)*channels + ch: out = istft(spec * mask) (complex multiply).
 * frame_pad     zero frames virtually added before and after (HTDemucs._ispec pads 2): frame index
 *               t of the padded sequence reads spec frame t - frame_pad.
 * out_start     untrimmed OLA position of output sample 0 (n_fft/2 for center=True, plus any crop).
 * out_len       samples produced per row.
 * weight        NULL or [out_len] device floats multiplied into the output (chunk window).
 * dst           out rows are written at
 *                 dst + (stem*channels + ch)*dst_ch_stride + chunk*dst_chunk_stride + place_c + p
 *               for p in [0, out_len) when 0 <= place_c + p < dst_limit, with
 *               place_c = dst_offsets[c] (device int64) if non-NULL else dst_off0 + c*dst_off_step.
 *               Dense chunk waves: dst_ch_stride = out_len, dst_chunk_stride = stems*channels*out_len,
 *               place = 0, dst_limit = out_len.
 */
int al_istft(const al_plan* plan, const float* spec, const float* mask, int layout, int n_bins_in,
             int n_frames_in, int frame_pad, int n_chunks, int stems, int channels, int spec_has_stems,
             int zero_low_bins, int out_start, int out_len, const float* weight, float* dst,
             int64_t dst_ch_stride, int64_t dst_chunk_stride, const int64_t* dst_offsets,
             int64_t dst_off0, int64_t dst_off_step, int64_t dst_limit, void* stream);

/*
 * K2b al_ola_gather -- deterministic windowed overlap-add of chunk outputs into a track:
 *   track[r, p] = (halo_in[r, p - p0] + sum_{c ascending, off_c <= p < off_c + len_c}
 *                  mult_c * W_c[p - off_c] * chunks[c, r, p - off_c]) / max(sum_c mult_c * W_c[..], eps)
 * Replaces: `result[..., s:e] += x * window; counter[..., s:e] += window; result / counter` of
 *   upstream MDXSeparator.demix / MDXCSeparator.demix / demucs.apply.apply_model (SURVEY.md
 *   A.1-A.3; stem_separator.py never sees it because it happens inside `separator.separate`,
 *   modules/separator/stem_separator.py:281).
 *
 * chunks        [n_chunks - data_chunk0, rows, chunk_len] fp32 dense chunk waves (al_istft dense output)
 *               of chunks data_chunk0 .. n_chunks-1.  Chunks before data_chunk0 are weight-only: they belong
 *               to the left neighbour rank, whose partial sums arrive in halo_in.
 * offsets       device int64 [n_chunks], ascending; len_c = min(chunk_len, n_total - off_c).
 * mult          device int32 [n_chunks] or NULL (all 1): tail-aligned chunks the reference evaluates
 *               several times.
 * wtab          device [n_tab, chunk_len] weight tables; chunk c uses table tab_id[c] (NULL -> 0);
 *               wtab NULL means weight 1.
 * [p0, p1)      track positions this call produces into track[r*track_stride + p] (chunk-range
 *               sharding owns a sub-range).  halo_in (NULL or [rows, p1-p0]) holds partial sums
 *               received from the left neighbour; when raw_out != 0 the un-normalised running sum
 *               is written instead (the partial sums a rank sends to its right neighbour).
 */
int al_ola_gather(const float* chunks, int n_chunks, int data_chunk0, int rows, int chunk_len, const int64_t* offsets,
                  const int32_t* mult, const float* wtab, const int32_t* tab_id, int64_t n_total,
                  int64_t p0, int64_t p1, const float* halo_in, int raw_out, float eps, float scale,
                  float* track, int64_t track_stride, void* stream);

/*
 * K3  al_resample_poly -- polyphase FIR resampler with scipy.signal.resample_poly semantics
 *   (zero-phase, zero-padded ends): out[m] = sum_k taps[k] * x_up[m*down + half - k].
 * Replaces: librosa.load(sr=44100) resampling (modules/separator/stem_separator.py:865) /
 *   res_type "polyphase" (modules/rvc/infer/lib/uvr5_pack/lib_v5/model_param_init.py:22).
 * taps = DEVICE pointer to n_taps floats (already multiplied by `up`), n_taps odd, half=(n_taps-1)/2.
 * in [rows, n_in] (row stride in_stride) -> out [rows, n_out] (row stride out_stride),
 * n_out = ceil(n_in*up/down).
 */
int al_resample_poly(const float* in, int64_t in_stride, float* out, int64_t out_stride, int rows,
                     int64_t n_in, int64_t n_out, int up, int down, const float* taps, int n_taps,
                     void* stream);

/*
 * al_sub_scaled -- secondary stem: out = a - b (elementwise), the `mix - primary` complement of
 * upstream MDXCSeparator / MDXSeparator when spectral inversion is off.
 */
int al_sub(const float* a, const float* b, float* out, int64_t n, void* stream);

/*
 * Row-wise operators of the RoFormer mask network's bf16 inference path.  Each replaces a chain of
 * elementwise PyTorch kernels that upstream bs_roformer / mel_band_roformer run between two dense
 * contractions (driven by MDXCSeparator.demix, reference call site
 * modules/separator/stem_separator.py:281).  All tensors are DEVICE bf16, row-major, 16-byte aligned.
 *
 * al_rmsnorm_bf16 -- upstream RMSNorm.forward: F.normalize(x, dim=-1) * sqrt(dim) * gamma.
 *   x [n_rows, dim]; if bias != NULL, x += bias is applied (and stored) first -- the deferred bias
 *   of the previous FeedForward's output projection.  out may alias x.  scale = sqrt(dim), eps 1e-12.
 * al_rotary_bf16 -- upstream rotary_embed.rotate_queries_or_keys on q and k, in place.
 *   q, k [n_rows, heads*dim_head]; cos_sin [pos_mod, dim_head/2, 2] fp32; the position of a row is
 *   (row / pos_div) % pos_mod (rows are tokens of a [batch, time, band] grid).
 * al_gate_sigmoid_bf16 -- upstream Attention: out * to_gates(x).sigmoid(), in place.
 *   o [n_rows, heads*dim_head], gates [n_rows, heads].
 * al_gelu_bf16 -- upstream FeedForward's nn.GELU() (exact, erf form) between its two Linear layers, in place.
 *   x [n] bf16, n a multiple of 8.
 * al_band_attention_bf16 -- upstream Attention.forward of the frequency (band-axis) transformer:
 *   softmax(q k^T * scale) v per (sequence, head); q, k, v, o [n_seq * seq_len, heads * 64] bf16, token (s, f) in row
 *   s * seq_len + f; seq_len <= 64, dim_head = 64.  gates (nullable) [n_seq * seq_len, heads] bf16 with row stride gate_ld
 *   (0 = heads): o *= sigmoid(gate);
 *   cos_sin (nullable) [seq_len, 32, 2] fp32: q and k are rotated by their band position first (al_rotary_bf16 semantics).  Opt-in on the host side (AUDIOLAB_B200_BAND_ATTN=1); the default
 *   calls the library attention (cuDNN through PyTorch).
 */
int al_rmsnorm_bf16(void* x, const float* gamma, const float* bias, void* out, int64_t n_rows, int dim, float scale,
                    float eps, void* stream);
int al_rotary_bf16(void* q, void* k, const float* cos_sin, int64_t n_rows, int heads, int dim_head, int64_t pos_div,
                   int pos_mod, void* stream);
int al_gate_sigmoid_bf16(void* o, const void* gates, int64_t n_rows, int heads, int dim_head, void* stream);
/* as al_gate_sigmoid_bf16 with a row stride: gates[row * gate_ld + h] (the gate columns of the fused to_qkv + to_gates GEMM);
 * fp16 = 1: o and gates are IEEE half (the fp16-operand mode of the network, see al_gemm_args.operand_fp16); likewise the
 * `fp16` argument of al_band_attention_bf16 (q, k, v, o, gates half; cos_sin must then be NULL). */
int al_gate_sigmoid_ld_bf16(void* o, const void* gates, int64_t gate_ld, int64_t n_rows, int heads, int dim_head, int fp16,
                            void* stream);
int al_gelu_bf16(void* x, int64_t n, void* stream);
int al_band_attention_bf16(const void* q, const void* k, const void* v, void* o, const void* gates, int64_t gate_ld,
                           const float* cos_sin, int64_t n_seq, int seq_len, int heads, int dim_head, float scale, int fp16,
                           void* stream);
/*
 * al_time_attention_bf16 -- upstream Attention.forward of the time-axis transformer (csrc/al_fattn.cu, tcgen05 flash
 * attention): softmax(q k^T * scale) v per (batch, inner index, head) over the seq_len positions of the sequence, then
 * o *= sigmoid(gate).  q, k, v, o [n_batch * seq_len * inner, heads * 64] 16-bit (bfloat16, or IEEE half with fp16 = 1),
 * token (b, t, i) in row (b * seq_len + t) * inner + i: for the RoFormer inner = number of bands, so consecutive positions
 * of one sequence are `inner` rows apart and no transposition copy is needed on either side (the reference path runs
 * rearrange 'b t f d -> (b f) t d' + F.scaled_dot_product_attention + gate multiply + rearrange back).
 * gates (nullable) [rows, heads] with row stride gate_ld (0 = heads).  dim_head = 64; any seq_len.
 */
int al_time_attention_bf16(const void* q, const void* k, const void* v, void* o, const void* gates, int64_t gate_ld,
                           int64_t n_batch, int seq_len, int inner, int heads, int dim_head, float scale, int fp16,
                           void* stream);

/*
 * al_gemm_bf16 -- K4: one nn.Linear (or a batch of `groups` of them) of the RoFormer mask network on the tcgen05
 * tensor cores, with the row-wise work upstream runs around it folded into the epilogue (csrc/al_gemm.cu).
 * Replaces, inside upstream BSRoformer / MelBandRoformer.forward (driven from the reference at
 * modules/separator/stem_separator.py:281 under autocast, :106): Attention.to_qkv / to_gates / to_out,
 * FeedForward's two Linear layers, the RMSNorm in front of them, rotary_embed.rotate_queries_or_keys, nn.GELU, the
 * residual adds, and the per-band Linear layers of BandSplit / MaskEstimator (groups > 1).
 *
 *   acc[g][m, n] = sum_k A[g][m, k] * W[g][n, k]      A [groups][M, K] bf16 (row stride lda, group stride
 *   a_group_stride, in elements), W [groups][N, K] bf16 (ldw, w_group_stride); fp32 accumulation.
 *
 * epi = AL_GEMM_EPI_BF16:  v = acc * rowscale[m] + bias[n];  rotary on columns < rot_cols;  act;  bf16 store.
 *   rowscale[m] = ss_scale / max(sqrt(sum_p row_ss[m * ss_parts + p]), ss_eps) if row_ss != NULL, else 1
 *     (RMSNorm(x) W^T = diag(rowscale) x (gamma (.) W)^T: the caller folds gamma into W once).
 *   bias fp32 [groups][N] or NULL.  cos_sin fp32 [pos_mod][32][2] or NULL: column pair (2i, 2i+1) of each 64-wide
 *     head turns by the angle of position (m / pos_div) % pos_mod (al_rotary_bf16 semantics); rot_cols % 64 == 0.
 *   act: AL_GEMM_ACT_NONE / _GELU (exact erf form) / _TANH.
 *   Output columns [i * out_split, (i + 1) * out_split) go to out[i] (row stride ldo[i], group stride
 *     o_group_stride[i]); out_split = 0 means one output.  At most 4 outputs.
 * epi = AL_GEMM_EPI_RESIDUAL (N % 128 == 0):  x32[m, n] += acc + bias[n] in place (fp32 residual stream),
 *   xb[m, n] = bf16(x32[m, n]),  ss_out[m * (N / S) + n / S] = sum over that S-column slab of x32[m, n]^2, with the
 *   slab width S = 128 if N % 256 == 0, else 64.
 *
 * epi = AL_GEMM_EPI_GLU (N % 16 == 0):  the caller interleaves W's rows as (a_0, b_0, a_1, b_1, ...) (and bias likewise);
 *   out[0][m, i] = (acc[m, 2i] + bias[2i]) * sigmoid(acc[m, 2i+1] + bias[2i+1]) in FP32, N / 2 columns (row stride ldo[0] in
 *   floats; out_split > 0 = number of output columns that exist, when W carries zero rows up to the multiple of 16) --
 *   upstream MaskEstimator's last Linear + nn.GLU, written straight into the mask tensor.
 *
 * K, lda, ldw, ldo, ldxb multiples of 8, ldx of 4, N of 8; all pointers 16-byte aligned.  max_ctas = 0 uses every SM.
 *
 * al_band_norm -- upstream BandSplit's per-band RMSNorm: for each band j, out[m, off_j : off_{j+1}] =
 *   bf16(F.normalize(x[m, off_j : off_{j+1}]) * sqrt(off_{j+1} - off_j) * gamma[off_j : off_{j+1}]); x fp32 (row stride ldx),
 *   out bf16 (IEEE half if out_fp16; row stride ldo), band_off DEVICE int32 [n_bands + 1].  The A operand of the grouped
 *   band-split GEMM.
 */
#define AL_GEMM_EPI_BF16 0
#define AL_GEMM_EPI_RESIDUAL 1
#define AL_GEMM_EPI_GLU 2
#define AL_GEMM_ACT_NONE 0
#define AL_GEMM_ACT_GELU 1
#define AL_GEMM_ACT_TANH 2

typedef struct al_gemm_args {
    const void* A;
    const void* W;
    int64_t M;
    int32_t N, K, groups;
    int64_t lda, a_group_stride, ldw, w_group_stride;
    int32_t epi, act;
    const float* bias;
    const float* row_ss;
    int32_t ss_parts;
    float ss_scale, ss_eps;
    const float* cos_sin;
    int64_t pos_div;
    int32_t pos_mod, rot_cols;
    void* out[4];
    int64_t ldo[4], o_group_stride[4];
    int32_t out_split;
    float* x32;
    void* xb;
    int64_t ldx, x_group_stride, ldxb, xb_group_stride;
    float* ss_out;
    int32_t max_ctas;
    int32_t no_accumulate;        /* EPI_RESIDUAL: 1 = x32 = acc + bias (start of the stream, x32 is not read) */
    int64_t side_row_stride;      /* 0 = default (row_ss / ss_out indexed by group * M + row); else the row index of the per-row */
    int64_t side_group_stride;    /*   side arrays is group * side_group_stride + row * side_row_stride (band-grouped calls)  */
    int32_t operand_fp16;         /* 1 = every 16-bit tensor of the call (A, W, out, xb) is IEEE half instead of bfloat16: same */
                                  /*   tensor-core rate, 11 instead of 8 significand bits, outputs saturate at +-65504          */
} al_gemm_args;

int al_gemm_bf16(const al_gemm_args* args, void* stream);
int al_band_norm(const float* x, int64_t ldx, const float* gamma, const int32_t* band_off, int n_bands, void* out, int64_t ldo,
                 int64_t n_rows, float eps, int out_fp16, void* stream);

/*
 * al_resid_prepare -- start (or re-normalise) the fp32 residual stream the residual epilogue of al_gemm_bf16 keeps:
 *   y = x_in (+ bias);  if gamma != NULL: y = F.normalize(y, dim=-1) * sqrt(dim) * gamma (upstream RMSNorm);
 *   x32 = y,  xb = bf16(y),  ss[m * ss_parts + p] = sum of y[m, p * dim / ss_parts ...)^2.
 * x_in fp32 [n_rows, dim] (may alias x32); dim a multiple of 8 * ss_parts, <= 2048; xb is IEEE half if xb_fp16.
 */
int al_resid_prepare(const float* x_in, const float* bias, const float* gamma, float* x32, void* xb, float* ss,
                     int64_t n_rows, int dim, int ss_parts, float eps, int xb_fp16, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AUDIOLAB_B200_H_ */
