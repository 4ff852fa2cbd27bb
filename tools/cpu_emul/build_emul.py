"""Builds tools/cpu_emul/_emul.so: the [emul-begin]..[emul-end] regions of the kernels that use only CTA / warp
barriers and shuffles (al_ola.cu, al_resample.cu, and the generic al_stft.cu / al_istft.cu with al_fft.cuh and the
generated 32-point FFT) compiled for the host on top of cuda_emul.h, so their addressing, masking, tiling and
arithmetic can be checked on a box without a GPU (tests/test_kernel_emulation.py).  Test infrastructure."""
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "..", "..", "audiolab_b200", "csrc")
SO = os.path.join(HERE, "_emul.so")

GLUE = r'''
using namespace al;

template <int D>
static int emul_stft_d(StftParams p, int rows) {
    const size_t smem = stft_tiling<D>(p);
    if (smem > sizeof(g_smem)) return -2;
    std::memset(g_smem, 0xFF, sizeof(g_smem));
    emul_launch(dim3(rows * p.tiles), dim3(Cfg<D>::UW * 32), [&] { stft_kernel<D>(p); });
    return 0;
}

extern "C" int emul_stft(int n_fft, int hop, const float* track, long long n_valid, long long ch_stride, int channels,
                         long long off0, long long off_step, int n_chunks, int chunk_len, int center, int n_frames,
                         const float* window, const float* tw, const float* ctw, float* spec, int layout,
                         int n_bins_out, int zero_low_bins) {
    StftParams p{};
    p.track = track; p.n_valid = n_valid; p.ch_stride = ch_stride; p.channels = channels; p.chunk_offsets = nullptr;
    p.off0 = off0; p.off_step = off_step; p.chunk_len = chunk_len; p.center = center; p.hop = hop; p.n_frames = n_frames;
    p.window = window; p.tw = reinterpret_cast<const float2*>(tw); p.ctw = reinterpret_cast<const float2*>(ctw);
    p.spec = spec; p.layout = layout; p.n_bins_out = n_bins_out; p.zero_low_bins = zero_low_bins;
    const int rows = n_chunks * channels;
    switch (n_fft) {
        case 2048: return emul_stft_d<2>(p, rows);
        case 4096: return emul_stft_d<4>(p, rows);
        case 6144: return emul_stft_d<6>(p, rows);
    }
    return -1;
}

template <int D>
static int emul_istft_d(IstftParams p, int n_chunks) {
    const int rows = n_chunks * p.stems * p.channels;
    const size_t smem = istft_tiling<D>(p, rows);
    if (smem > sizeof(g_smem)) return -2;
    std::memset(g_smem, 0xFF, sizeof(g_smem));
    emul_launch(dim3(rows * p.segs), dim3(Cfg<D>::UW * 32), [&] { istft_kernel<D>(p); });
    return p.segs;
}

extern "C" int emul_istft(int n_fft, int hop, const float* spec, const float* mask, int layout, int n_bins_in,
                          int n_frames_in, int frame_pad, int n_chunks, int stems, int channels, int spec_has_stems,
                          int zero_low_bins, const float* window, const float* tw, const float* ctw,
                          const float* inv_env, int out_start, int out_len, const float* weight, float* dst,
                          long long dst_ch_stride, long long dst_chunk_stride, long long dst_off0,
                          long long dst_off_step, long long dst_limit) {
    IstftParams p{};
    p.spec = spec; p.mask = mask; p.layout = layout; p.n_bins_in = n_bins_in; p.n_frames_in = n_frames_in;
    p.frame_pad = frame_pad; p.n_frames_total = n_frames_in + 2 * frame_pad; p.stems = stems; p.channels = channels;
    p.spec_has_stems = spec_has_stems; p.zero_low_bins = zero_low_bins; p.hop = hop; p.window = window;
    p.tw = reinterpret_cast<const float2*>(tw); p.ctw = reinterpret_cast<const float2*>(ctw); p.inv_env = inv_env;
    p.out_start = out_start; p.out_len = out_len; p.weight = weight; p.dst = dst; p.dst_ch_stride = dst_ch_stride;
    p.dst_chunk_stride = dst_chunk_stride; p.dst_offsets = nullptr; p.dst_off0 = dst_off0; p.dst_off_step = dst_off_step;
    p.dst_limit = dst_limit;
    switch (n_fft) {
        case 2048: return emul_istft_d<2>(p, n_chunks);
        case 4096: return emul_istft_d<4>(p, n_chunks);
        case 6144: return emul_istft_d<6>(p, n_chunks);
    }
    return -1;
}

extern "C" void emul_rmsnorm(void* x, const float* gamma, const float* bias, void* out, long long n_rows, int dim, float scale,
                             float eps) {
    const unsigned grid = (unsigned)((n_rows + 7) / 8);
    auto* xb = reinterpret_cast<__nv_bfloat16*>(x);
    auto* ob = reinterpret_cast<__nv_bfloat16*>(out);
    if (dim <= 512) emul_launch(dim3(grid), dim3(256), [&] { rmsnorm_bf16_kernel<2>(xb, gamma, bias, ob, n_rows, dim, scale, eps); });
    else if (dim <= 1024) emul_launch(dim3(grid), dim3(256), [&] { rmsnorm_bf16_kernel<4>(xb, gamma, bias, ob, n_rows, dim, scale, eps); });
    else emul_launch(dim3(grid), dim3(256), [&] { rmsnorm_bf16_kernel<8>(xb, gamma, bias, ob, n_rows, dim, scale, eps); });
}

extern "C" void emul_rotary(void* q, void* k, const float* cs, long long n_rows, int heads, int dim_head, long long pos_div,
                            int pos_mod) {
    const int vec_per_row = heads * dim_head / 8;
    const long long n_vec = n_rows * vec_per_row;
    emul_launch(dim3((unsigned)((n_vec + 255) / 256)), dim3(256), [&] {
        rotary_bf16_kernel(reinterpret_cast<uint4*>(q), reinterpret_cast<uint4*>(k), reinterpret_cast<const float2*>(cs), n_vec,
                           vec_per_row, dim_head, pos_div, pos_mod);
    });
}

extern "C" void emul_gate(void* o, const void* gates, long long n_rows, int heads, int dim_head) {
    const int vec_per_row = heads * dim_head / 8;
    const long long n_vec = n_rows * vec_per_row;
    emul_launch(dim3((unsigned)((n_vec + 255) / 256)), dim3(256), [&] {
        gate_h16_kernel<false>(reinterpret_cast<uint4*>(o), reinterpret_cast<const __nv_bfloat16*>(gates), n_vec, vec_per_row, heads, dim_head);
    });
}

extern "C" void emul_band_attn(const void* q, const void* k, const void* v, void* o, const void* gates, const float* cos_sin,
                               int n_seq, int F, int H, float scale) {
    emul_launch(dim3(n_seq * H), dim3(128), [&] {
        band_attn_bf16_kernel<false>(reinterpret_cast<const __nv_bfloat16*>(q), reinterpret_cast<const __nv_bfloat16*>(k),
                              reinterpret_cast<const __nv_bfloat16*>(v), reinterpret_cast<__nv_bfloat16*>(o),
                              reinterpret_cast<const __nv_bfloat16*>(gates), reinterpret_cast<const float2*>(cos_sin), F, H, scale, 0);
    });
}

extern "C" void emul_gelu(void* x, long long n, int grid_x) {
    static std::vector<unsigned short> lut(65536);
    emul_launch(dim3(256), dim3(256), [&] { gelu_lut_init_kernel(lut.data()); });
    const long long n_vec = n / 8;
    emul_launch(dim3(grid_x), dim3(kGeluThreads), [&] {
        gelu_bf16_kernel(reinterpret_cast<uint4*>(x), n_vec, reinterpret_cast<const uint4*>(lut.data()));
    });
}

extern "C" int emul_stft_pk(const float* track, long long n_valid, long long ch_stride, long long off0, long long off_step,
                            int n_chunks, int chunk_len, int center, int hop, int n_frames, const float* window,
                            const float* tw, const float* ctw_half, float* spec, int layout, int n_bins_out,
                            int zero_low_bins, int aligned, int grid_x) {
    StftPkParams p{};
    p.track = track; p.n_valid = n_valid; p.ch_stride = ch_stride; p.chunk_offsets = nullptr; p.off0 = off0;
    p.off_step = off_step; p.n_chunks = n_chunks; p.chunk_len = chunk_len; p.center = center; p.hop = hop;
    p.n_frames = n_frames; p.window = window; p.tw = reinterpret_cast<const float2*>(tw);
    p.ctw_half = reinterpret_cast<const float2*>(ctw_half); p.spec = spec; p.layout = layout; p.n_bins_out = n_bins_out;
    p.zero_low_bins = zero_low_bins; p.aligned = aligned;
    const size_t smem = pk_launch_shape(p);
    if (smem == 0 || smem > sizeof(g_smem)) return -2;
    std::memset(g_smem, 0xFF, sizeof(g_smem));
    const bool full = n_bins_out == 1025 && zero_low_bins == 0;
    const dim3 grid(std::min(p.total_tiles, grid_x)), block(kPkThreads);
    if (layout == 3) {
        if (full) emul_launch(grid, block, [&] { stft_pk2_kernel<3, true>(p); });
        else emul_launch(grid, block, [&] { stft_pk2_kernel<3, false>(p); });
    } else {
        if (full) emul_launch(grid, block, [&] { stft_pk2_kernel<0, true>(p); });
        else emul_launch(grid, block, [&] { stft_pk2_kernel<0, false>(p); });
    }
    return p.n_stages;
}

extern "C" int emul_istft_pk(const float* spec, const float* mask, int n_frames, int stems, int spec_has_stems, int hop,
                             const float* window, const float* tw, const float* ctw_full, const float* inv_env,
                             int out_start, int out_len, const float* weight, float* dst, long long dst_ch_stride,
                             long long dst_chunk_stride, long long dst_off0, long long dst_off_step, long long dst_limit,
                             int n_chunks, int warps, int n_sm, int pre, int ola_fast) {
    IstftPkParams p{};
    p.spec = reinterpret_cast<const float4*>(spec); p.mask = reinterpret_cast<const float4*>(mask); p.n_frames = n_frames;
    p.stems = stems; p.spec_has_stems = spec_has_stems; p.hop = hop; p.window = window;
    p.tw = reinterpret_cast<const float2*>(tw); p.ctw = reinterpret_cast<const float2*>(ctw_full); p.inv_env = inv_env;
    p.out_start = out_start; p.out_len = out_len; p.weight = weight; p.dst = dst; p.dst_ch_stride = dst_ch_stride;
    p.dst_chunk_stride = dst_chunk_stride; p.dst_offsets = nullptr; p.dst_off0 = dst_off0; p.dst_off_step = dst_off_step;
    p.dst_limit = dst_limit;
    p.ola_fast = ola_fast;
    const size_t smem = ip_launch_shape(p, n_chunks, n_sm, warps);
    if (smem > sizeof(g_smem)) return -2;
    std::memset(g_smem, 0xFF, sizeof(g_smem));
    const dim3 grid(n_chunks * stems * p.segs), block(warps * 32);
#define EMUL_IP(WW, PP)                                                                          \
    do {                                                                                         \
        if (mask) emul_launch(grid, block, [&] { istft_pk2_kernel<true, WW, PP>(p); });         \
        else emul_launch(grid, block, [&] { istft_pk2_kernel<false, WW, PP>(p); });             \
    } while (0)
    if (warps == 8) { if (pre == 0) EMUL_IP(8, 0); else EMUL_IP(8, 2); }
    else { if (pre == 0) EMUL_IP(4, 0); else if (pre == 3) EMUL_IP(4, 3); else EMUL_IP(4, 2); }
#undef EMUL_IP
    return p.segs;
}

extern "C" int emul_istft_pk4(const float* spec, const float* mask, int n_frames, int stems, int spec_has_stems, int hop,
                              const float* window, const float* tw, const float* ctw_full, const float* inv_env,
                              int out_start, int out_len, const float* weight, float* dst, long long dst_ch_stride,
                              long long dst_chunk_stride, long long dst_off0, long long dst_off_step, long long dst_limit,
                              int n_chunks, int consumers, int n_sm) {
    IstftPkParams p{};
    p.spec = reinterpret_cast<const float4*>(spec); p.mask = reinterpret_cast<const float4*>(mask); p.n_frames = n_frames;
    p.stems = stems; p.spec_has_stems = spec_has_stems; p.hop = hop; p.window = window;
    p.tw = reinterpret_cast<const float2*>(tw); p.ctw = reinterpret_cast<const float2*>(ctw_full); p.inv_env = inv_env;
    p.out_start = out_start; p.out_len = out_len; p.weight = weight; p.dst = dst; p.dst_ch_stride = dst_ch_stride;
    p.dst_chunk_stride = dst_chunk_stride; p.dst_offsets = nullptr; p.dst_off0 = dst_off0; p.dst_off_step = dst_off_step;
    p.dst_limit = dst_limit;
    if (consumers == 5) {
        const size_t smem5 = sr_launch_shape(p, n_chunks, n_sm);
        if (smem5 > sizeof(g_smem)) return -2;
        std::memset(g_smem, 0xFF, sizeof(g_smem));
        const dim3 grid5(n_chunks * stems * p.segs), block5(kSrThreads);
        if (mask) emul_launch(grid5, block5, [&] { istft_pk5_kernel<true>(p); });
        else emul_launch(grid5, block5, [&] { istft_pk5_kernel<false>(p); });
        return p.segs;
    }
    const size_t smem = tk_launch_shape(p, n_chunks, n_sm);
    if (smem > sizeof(g_smem)) return -2;
    std::memset(g_smem, 0xFF, sizeof(g_smem));
    const dim3 grid(n_chunks * stems * p.segs), block(kTkThreads);
    if (mask) emul_launch(grid, block, [&] { istft_pk4_kernel<true>(p); });
    else emul_launch(grid, block, [&] { istft_pk4_kernel<false>(p); });
    return p.segs;
}

extern "C" void emul_ola_gather(const float* chunks, int n_chunks, int data_chunk0, int rows, int chunk_len,
                                const long long* offsets, const int* mult, const float* wtab, const int* tab_id,
                                long long n_total, long long p0, long long p1, const float* halo_in, int raw_out,
                                float eps, float scale, float* track, long long track_stride, int block) {
    const long long per_cta = (long long)block * kOlaVec;
    if (rows % 2 == 0) {
        dim3 grid((unsigned)((p1 - p0 + per_cta - 1) / per_cta), (unsigned)std::min(rows / 2, 2));
        emul_launch(grid, dim3(block), [&] {
            ola_gather_kernel<2>(chunks, n_chunks, data_chunk0, rows, chunk_len, offsets, mult, wtab, tab_id, n_total, p0, p1,
                                 halo_in, raw_out, eps, scale, track, track_stride);
        });
    } else {
        dim3 grid((unsigned)((p1 - p0 + per_cta - 1) / per_cta), (unsigned)std::min(rows, 2));
        emul_launch(grid, dim3(block), [&] {
            ola_gather_kernel<1>(chunks, n_chunks, data_chunk0, rows, chunk_len, offsets, mult, wtab, tab_id, n_total, p0, p1,
                                 halo_in, raw_out, eps, scale, track, track_stride);
        });
    }
}

extern "C" int emul_resample(const float* in, long long in_stride, float* out, long long out_stride, int rows,
                             long long n_in, long long n_out, int up, int down, const float* taps, int n_taps,
                             int grid_x) {
    ResampleRbParams q{};
    q.in = in; q.in_stride = in_stride; q.out = out; q.out_stride = out_stride;
    q.n_in = n_in; q.n_out = n_out; q.up = up; q.down = down; q.taps = taps; q.n_taps = n_taps;
    size_t smem = 0;
    if (!resample_rb_plan(q, rows, smem)) return -1;
    if (smem > sizeof(g_smem)) return -2;
    std::memset(g_smem, 0xFF, sizeof(g_smem));   // NaNs: reading an unstaged word shows up in the result
    const unsigned grid = (unsigned)std::min<long long>(q.total_tiles, grid_x);
    emul_launch(dim3(grid), dim3(kRbThreads), [&] { resample_rb_kernel(q); });
    return q.vec_in * 2 + q.vec_out;
}
'''


def region(path):
    s = open(path).read()
    return "\n".join(re.findall(r"// \[emul-begin\]\n(.*?)// \[emul-end\]", s, flags=re.S))


def build(force=False):
    srcs = [os.path.join(CSRC, f) for f in ("al_kernels.h", "al_ola.cu", "al_resample.cu", "al_stft.cu", "al_istft.cu",
                                            "al_istft_pk.cu", "al_stft_pk.cu", "al_netops.cu", "al_attn.cu")]
    whole = [os.path.join(CSRC, f) for f in ("fft32_gen.cuh", "al_fft.cuh", "fft32p_gen.cuh", "al_fftp.cuh")]   # taken whole
    deps = srcs + whole + [os.path.join(HERE, "cuda_emul.h"), __file__]
    if not force and os.path.exists(SO) and all(os.path.getmtime(SO) > os.path.getmtime(d) for d in deps):
        return SO
    gen = os.path.join(HERE, "_emul_gen.cpp")
    with open(gen, "w") as f:
        f.write('#include "cuda_emul.h"\n')
        for w in whole:
            body = re.sub(r'^#(pragma once|include .*)$', "", open(w).read(), flags=re.M)
            f.write(f"// ---- {os.path.basename(w)}\n" + body + "\n")
        f.write("namespace al {\n")
        for s in srcs:
            f.write(f"// ---- from {os.path.basename(s)}\n" + region(s) + "\n")
        f.write("}  // namespace al\n")
        f.write(GLUE)
    subprocess.run(["g++", "-std=c++20", "-O1", "-DAL_CPU_EMUL", "-g", "-mfma", "-ffp-contract=fast", "-shared", "-fPIC", "-pthread",
                    "-I", HERE, "-o", SO, gen], check=True)
    return SO


if __name__ == "__main__":
    print(build(force=True))
