// C ABI of libaudiolab_b200.so -- see include/audiolab_b200.h for the contract of every entry point
// and the reference interface each one replaces.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/audiolab_b200.h"
#include "al_gemm.h"
#include "al_kernels.h"

namespace al {
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
int sm_count() {
    static std::atomic<int> cache[64];
    int d = 0;
    cudaGetDevice(&d);
    int n = cache[d & 63].load(std::memory_order_relaxed);
    if (n == 0) {
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d);
        if (n <= 0) n = 148;
        cache[d & 63].store(n, std::memory_order_relaxed);
    }
    return n;
}
}  // namespace al

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

int cuda_fail(cudaError_t e, const char* what) {
    return fail(AL_E_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

}  // namespace

struct al_plan {
    int n_fft, hop, normalized, D;
    float* d_win_a = nullptr;     // analysis window * (normalized ? n_fft^-1/2 : 1)
    float* d_win_s = nullptr;     // synthesis window * 1/n_fft * (normalized ? n_fft^1/2 : 1)
    float* d_win_raw = nullptr;   // window as given (for the OLA envelope)
    float2* d_tw = nullptr;       // [32*32]
    float2* d_ctw = nullptr;      // [(D-1)*513]
    float2* d_ctw_full = nullptr; // n_fft 2048 only: [1024] exp(-2 pi i k / 2048) (packed stereo fast path of K2)
    float2* d_ctw_half = nullptr; // n_fft 2048 only: [544] 0.5 * exp(-2 pi i k / 2048) (packed stereo fast path)
    struct Env {
        float* table = nullptr;
        cudaEvent_t ready = nullptr;   // recorded on `stream` after the build
        cudaStream_t stream = nullptr;
        unsigned long long stamp = 0;  // LRU
    };
    static constexpr int kMaxEnv = 8;
    std::mutex mu;
    std::map<int, Env> env;       // n_frames_total -> 1 / sum(w^2) table (bounded LRU)
    unsigned long long env_clock = 0;
    int device = -1;              // device the tables live on
};

// A plan's tables live on the device that was current when it was created; launching with another current device
// would hand its kernels foreign pointers.
static int check_device(const al_plan* plan, const char* what) {
    int dev = -1;
    cudaGetDevice(&dev);
    if (dev != plan->device)
        return fail(AL_E_ARG, "%s: the plan was created on device %d but device %d is current", what, plan->device, dev);
    return AL_OK;
}

extern "C" {

int al_version(void) { return 110; }

const char* al_last_error(void) { return g_err.c_str(); }

int64_t al_launch_count(void) { return al::g_launches.load(std::memory_order_relaxed); }

int al_plan_create(int n_fft, int hop, const float* window_host, int normalized, al_plan** out) {
    if (!out) return fail(AL_E_ARG, "al_plan_create: out is NULL");
    *out = nullptr;
    if (n_fft != 2048 && n_fft != 4096 && n_fft != 6144)
        return fail(AL_E_UNSUPPORTED, "al_plan_create: n_fft %d not in {2048, 4096, 6144}", n_fft);
    if (hop <= 0 || hop > n_fft) return fail(AL_E_ARG, "al_plan_create: bad hop %d", hop);
    al_plan* p = new al_plan;
    p->n_fft = n_fft;
    p->hop = hop;
    p->normalized = normalized ? 1 : 0;
    p->D = n_fft / 1024;
    cudaGetDevice(&p->device);
    const int N = n_fft, D = p->D;
    const double kPi = 3.14159265358979323846;
    std::vector<float> raw(N), wa(N), ws(N);
    for (int i = 0; i < N; ++i) {
        // torch.hann_window(N, periodic=True) evaluated in float32 like torch does is within 1 ulp of this
        raw[i] = window_host ? window_host[i] : (float)(0.5 - 0.5 * cos(2.0 * kPi * i / N));
        const double sa = normalized ? 1.0 / sqrt((double)N) : 1.0;
        const double ss = (normalized ? sqrt((double)N) : 1.0) / (double)N;
        wa[i] = (float)((double)raw[i] * sa);
        ws[i] = (float)((double)raw[i] * ss);
    }
    std::vector<float2> tw(1024), ctw((size_t)(D - 1) * 513);
    for (int k1 = 0; k1 < 32; ++k1)
        for (int n2 = 0; n2 < 32; ++n2) {
            const double a = -2.0 * kPi * (double)(k1 * n2) / 1024.0;
            tw[k1 * 32 + n2] = make_float2((float)cos(a), (float)sin(a));
        }
    for (int r = 1; r < D; ++r)
        for (int k = 0; k <= 512; ++k) {
            const double a = -2.0 * kPi * (double)r * (double)k / (double)N;
            ctw[(size_t)(r - 1) * 513 + k] = make_float2((float)cos(a), (float)sin(a));
        }
    cudaError_t e;
#define AL_UP(dst, src, bytes)                                                     \
    do {                                                                           \
        e = cudaMalloc((void**)&(dst), (bytes));                                   \
        if (e == cudaSuccess) e = cudaMemcpy((dst), (src), (bytes), cudaMemcpyHostToDevice); \
        if (e != cudaSuccess) { al_plan_destroy(p); return cuda_fail(e, "al_plan_create"); } \
    } while (0)
    AL_UP(p->d_win_a, wa.data(), N * sizeof(float));
    AL_UP(p->d_win_s, ws.data(), N * sizeof(float));
    AL_UP(p->d_win_raw, raw.data(), N * sizeof(float));
    AL_UP(p->d_tw, tw.data(), tw.size() * sizeof(float2));
    AL_UP(p->d_ctw, ctw.data(), ctw.size() * sizeof(float2));
    if (D == 2) {
        std::vector<float2> half(544, make_float2(0.f, 0.f));
        for (int k = 0; k <= 512; ++k) {
            const double a = -2.0 * kPi * (double)k / (double)N;
            half[k] = make_float2((float)(0.5 * cos(a)), (float)(0.5 * sin(a)));
        }
        AL_UP(p->d_ctw_half, half.data(), half.size() * sizeof(float2));
        std::vector<float2> full(1024);
        for (int k = 0; k < 1024; ++k) {
            const double a = -2.0 * kPi * (double)k / (double)N;
            full[k] = make_float2((float)cos(a), (float)sin(a));
        }
        AL_UP(p->d_ctw_full, full.data(), full.size() * sizeof(float2));
    }
#undef AL_UP
    *out = p;
    return AL_OK;
}

int al_plan_destroy(al_plan* p) {
    if (!p) return AL_OK;
    cudaFree(p->d_win_a);
    cudaFree(p->d_win_s);
    cudaFree(p->d_win_raw);
    cudaFree(p->d_tw);
    cudaFree(p->d_ctw);
    cudaFree(p->d_ctw_half);
    cudaFree(p->d_ctw_full);
    for (auto& kv : p->env) {
        cudaFree(kv.second.table);
        if (kv.second.ready) cudaEventDestroy(kv.second.ready);
    }
    delete p;
    return AL_OK;
}

int al_stft(const al_plan* plan, const float* track, int64_t n_valid, int64_t ch_stride, int channels,
            const int64_t* chunk_offsets, int64_t off0, int64_t off_step, int n_chunks, int chunk_len,
            int center_pad, int n_frames, float* spec, int layout, int n_bins_out, int zero_low_bins,
            void* stream) {
    if (!plan || !track || !spec) return fail(AL_E_ARG, "al_stft: NULL argument");
    if (n_chunks == 0 || n_frames == 0) return AL_OK;
    if (int rc = check_device(plan, "al_stft")) return rc;
    if (channels <= 0 || n_chunks < 0 || n_frames < 0 || chunk_len <= 0)
        return fail(AL_E_ARG, "al_stft: bad sizes (channels %d chunks %d frames %d chunk_len %d)", channels,
                    n_chunks, n_frames, chunk_len);
    if (layout < 0 || layout > 3) return fail(AL_E_ARG, "al_stft: bad layout %d", layout);
    if (n_bins_out <= 0 || n_bins_out > plan->n_fft / 2 + 1) return fail(AL_E_ARG, "al_stft: bad n_bins_out %d", n_bins_out);
    if (center_pad < 0 || center_pad >= chunk_len)
        return fail(AL_E_ARG, "al_stft: center_pad %d must be < chunk_len %d (single reflection)", center_pad, chunk_len);
    // right edge: the last frame may reach at most chunk_len - 1 samples past the end
    const long long last = (long long)(n_frames - 1) * plan->hop - center_pad + plan->n_fft - 1;
    if (last > 2LL * (chunk_len - 1))
        return fail(AL_E_ARG, "al_stft: frames reach %lld, beyond a single reflection of chunk_len %d", last, chunk_len);
    if ((reinterpret_cast<uintptr_t>(spec) & 7) != 0) return fail(AL_E_ARG, "al_stft: spec must be 8-byte aligned");
    static const bool force_generic = getenv("AL_FORCE_GENERIC") != nullptr;
    if (!force_generic && plan->D == 2 && channels == 2 && (layout == 0 || layout == 3) &&
        (layout == 0 || (reinterpret_cast<uintptr_t>(spec) & 15) == 0)) {
        al::StftPkParams q{};
        q.track = track;
        q.n_valid = n_valid;
        q.ch_stride = ch_stride;
        q.chunk_offsets = reinterpret_cast<const long long*>(chunk_offsets);
        q.off0 = off0;
        q.off_step = off_step;
        q.n_chunks = n_chunks;
        q.chunk_len = chunk_len;
        q.center = center_pad;
        q.hop = plan->hop;
        q.n_frames = n_frames;
        q.window = plan->d_win_a;
        q.tw = plan->d_tw;
        q.ctw_half = plan->d_ctw_half;
        q.spec = spec;
        q.layout = layout;
        q.n_bins_out = n_bins_out;
        q.zero_low_bins = zero_low_bins;
        q.aligned = ((reinterpret_cast<uintptr_t>(track) & 15) == 0 && (ch_stride & 3) == 0) ? 1 : 0;
        cudaError_t e = al::launch_stft_pk(q, (cudaStream_t)stream);
        if (e != cudaSuccess) return cuda_fail(e, "al_stft (packed stereo path)");
        return AL_OK;
    }
    al::StftParams p{};
    p.track = track;
    p.n_valid = n_valid;
    p.ch_stride = ch_stride;
    p.channels = channels;
    p.chunk_offsets = reinterpret_cast<const long long*>(chunk_offsets);
    p.off0 = off0;
    p.off_step = off_step;
    p.chunk_len = chunk_len;
    p.center = center_pad;
    p.hop = plan->hop;
    p.n_frames = n_frames;
    p.window = plan->d_win_a;
    p.tw = plan->d_tw;
    p.ctw = plan->d_ctw;
    p.spec = spec;
    p.layout = layout;
    p.n_bins_out = n_bins_out;
    p.zero_low_bins = zero_low_bins;
    cudaError_t e = al::launch_stft(p, plan->n_fft, n_chunks * channels, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "al_stft");
    return AL_OK;
}

// 1 / sum(w^2) tables, one per frame count, built on the CALLER's stream (no host synchronisation) and published to
// other streams through an event.  The cache is bounded: a long-running process that sees a new MDX track length per
// song (invert_stem's whole-track iSTFT) would otherwise grow by ~4 bytes per sample per distinct length.
static int get_env(al_plan* plan, int n_frames_total, cudaStream_t stream, const float** out) {
    std::lock_guard<std::mutex> lk(plan->mu);
    auto it = plan->env.find(n_frames_total);
    if (it != plan->env.end()) {
        it->second.stamp = ++plan->env_clock;
        if (it->second.stream != stream) {
            cudaError_t e = cudaStreamWaitEvent(stream, it->second.ready, 0);
            if (e != cudaSuccess) return cuda_fail(e, "al_istft: envelope wait");
        }
        *out = it->second.table;
        return AL_OK;
    }
    while (plan->env.size() >= (size_t)al_plan::kMaxEnv) {
        auto victim = plan->env.begin();
        for (auto j = plan->env.begin(); j != plan->env.end(); ++j)
            if (j->second.stamp < victim->second.stamp) victim = j;
        cudaFree(victim->second.table);           // synchronises with every kernel that may still read the table
        cudaEventDestroy(victim->second.ready);
        plan->env.erase(victim);
    }
    const long long total = (long long)(n_frames_total - 1) * plan->hop + plan->n_fft;
    al_plan::Env ent;
    cudaError_t e = cudaMalloc((void**)&ent.table, total * sizeof(float));
    if (e != cudaSuccess) return cuda_fail(e, "al_istft: envelope alloc");
    e = cudaEventCreateWithFlags(&ent.ready, cudaEventDisableTiming);
    if (e == cudaSuccess) e = al::launch_env(plan->d_win_raw, plan->n_fft, plan->hop, n_frames_total, ent.table, stream);
    if (e == cudaSuccess) e = cudaEventRecord(ent.ready, stream);
    if (e != cudaSuccess) {
        cudaFree(ent.table);
        if (ent.ready) cudaEventDestroy(ent.ready);
        return cuda_fail(e, "al_istft: envelope build");
    }
    ent.stream = stream;
    ent.stamp = ++plan->env_clock;
    plan->env[n_frames_total] = ent;
    *out = ent.table;
    return AL_OK;
}

int al_istft(const al_plan* plan_c, const float* spec, const float* mask, int layout, int n_bins_in,
             int n_frames_in, int frame_pad, int n_chunks, int stems, int channels, int spec_has_stems,
             int zero_low_bins, int out_start, int out_len, const float* weight, float* dst,
             int64_t dst_ch_stride, int64_t dst_chunk_stride, const int64_t* dst_offsets,
             int64_t dst_off0, int64_t dst_off_step, int64_t dst_limit, void* stream) {
    al_plan* plan = const_cast<al_plan*>(plan_c);
    if (!plan || !spec || !dst) return fail(AL_E_ARG, "al_istft: NULL argument");
    if (n_chunks == 0 || out_len == 0) return AL_OK;
    if (int rc = check_device(plan, "al_istft")) return rc;
    if (n_chunks < 0 || stems <= 0 || channels <= 0 || n_frames_in <= 0 || out_len < 0 || frame_pad < 0)
        return fail(AL_E_ARG, "al_istft: bad sizes");
    if (layout < 0 || layout > 3) return fail(AL_E_ARG, "al_istft: bad layout %d", layout);
    if (mask && layout == 2) return fail(AL_E_UNSUPPORTED, "al_istft: mask multiply needs a complex layout");
    if (n_bins_in <= 0 || n_bins_in > plan->n_fft / 2 + 1) return fail(AL_E_ARG, "al_istft: bad n_bins_in %d", n_bins_in);
    const int T = n_frames_in + 2 * frame_pad;
    const long long ola_len = (long long)(T - 1) * plan->hop + plan->n_fft;
    if (out_start < 0 || (long long)out_start + out_len > ola_len)
        return fail(AL_E_ARG, "al_istft: [out_start %d, +%d) exceeds the overlap-add length %lld", out_start, out_len, ola_len);
    const float* inv_env = nullptr;
    int rc = get_env(plan, T, (cudaStream_t)stream, &inv_env);
    if (rc != AL_OK) return rc;
    static const bool force_generic = getenv("AL_FORCE_GENERIC") != nullptr;
    if (!force_generic && plan->D == 2 && channels == 2 && layout == 3 && n_bins_in == 1025 && zero_low_bins == 0 &&
        frame_pad == 0 && (reinterpret_cast<uintptr_t>(spec) & 15) == 0 && (reinterpret_cast<uintptr_t>(mask) & 15) == 0) {
        al::IstftPkParams q{};
        q.spec = reinterpret_cast<const float4*>(spec);
        q.mask = reinterpret_cast<const float4*>(mask);
        q.n_frames = n_frames_in;
        q.stems = stems;
        q.spec_has_stems = spec_has_stems;
        q.hop = plan->hop;
        q.window = plan->d_win_s;
        q.tw = plan->d_tw;
        q.ctw = plan->d_ctw_full;
        q.inv_env = inv_env;
        q.out_start = out_start;
        q.out_len = out_len;
        q.weight = weight;
        q.dst = dst;
        q.dst_ch_stride = dst_ch_stride;
        q.dst_chunk_stride = dst_chunk_stride;
        q.dst_offsets = reinterpret_cast<const long long*>(dst_offsets);
        q.dst_off0 = dst_off0;
        q.dst_off_step = dst_off_step;
        q.dst_limit = dst_limit;
        cudaError_t e = al::launch_istft_pk(q, n_chunks, (cudaStream_t)stream);
        if (e != cudaSuccess) return cuda_fail(e, "al_istft (packed stereo path)");
        return AL_OK;
    }
    al::IstftParams p{};
    p.spec = spec;
    p.mask = mask;
    p.layout = layout;
    p.n_bins_in = n_bins_in;
    p.n_frames_in = n_frames_in;
    p.frame_pad = frame_pad;
    p.n_frames_total = T;
    p.stems = stems;
    p.channels = channels;
    p.spec_has_stems = spec_has_stems;
    p.zero_low_bins = zero_low_bins;
    p.hop = plan->hop;
    p.window = plan->d_win_s;
    p.tw = plan->d_tw;
    p.ctw = plan->d_ctw;
    p.inv_env = inv_env;
    p.out_start = out_start;
    p.out_len = out_len;
    p.weight = weight;
    p.dst = dst;
    p.dst_ch_stride = dst_ch_stride;
    p.dst_chunk_stride = dst_chunk_stride;
    p.dst_offsets = reinterpret_cast<const long long*>(dst_offsets);
    p.dst_off0 = dst_off0;
    p.dst_off_step = dst_off_step;
    p.dst_limit = dst_limit;
    cudaError_t e = al::launch_istft(p, plan->n_fft, n_chunks, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "al_istft");
    return AL_OK;
}

int al_ola_gather(const float* chunks, int n_chunks, int data_chunk0, int rows, int chunk_len, const int64_t* offsets,
                  const int32_t* mult, const float* wtab, const int32_t* tab_id, int64_t n_total,
                  int64_t p0, int64_t p1, const float* halo_in, int raw_out, float eps, float scale,
                  float* track, int64_t track_stride, void* stream) {
    if (!chunks || !offsets || !track) return fail(AL_E_ARG, "al_ola_gather: NULL argument");
    if (n_chunks <= 0 || rows <= 0 || chunk_len <= 0 || p0 < 0 || p1 < p0 || data_chunk0 < 0 || data_chunk0 > n_chunks)
        return fail(AL_E_ARG, "al_ola_gather: bad sizes");
    cudaError_t e = al::launch_ola_gather(chunks, n_chunks, data_chunk0, rows, chunk_len,
                                          reinterpret_cast<const long long*>(offsets), mult, wtab, tab_id, n_total,
                                          p0, p1, halo_in, raw_out, eps, scale, track, track_stride,
                                          (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "al_ola_gather");
    return AL_OK;
}

int al_resample_poly(const float* in, int64_t in_stride, float* out, int64_t out_stride, int rows,
                     int64_t n_in, int64_t n_out, int up, int down, const float* taps, int n_taps,
                     void* stream) {
    if (!in || !out || !taps) return fail(AL_E_ARG, "al_resample_poly: NULL argument");
    if (rows < 0 || n_in < 0 || n_out < 0 || up <= 0 || down <= 0 || n_taps <= 0 || (n_taps & 1) == 0)
        return fail(AL_E_ARG, "al_resample_poly: bad sizes (n_taps must be odd)");
    if (n_out > (n_in * up + down - 1) / down)
        return fail(AL_E_ARG, "al_resample_poly: n_out %lld > ceil(n_in*up/down)", (long long)n_out);
    cudaError_t e = al::launch_resample(in, in_stride, out, out_stride, rows, n_in, n_out, up, down, taps, n_taps,
                                        (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "al_resample_poly");
    return AL_OK;
}

int al_sub(const float* a, const float* b, float* out, int64_t n, void* stream) {
    if (!a || !b || !out) return fail(AL_E_ARG, "al_sub: NULL argument");
    cudaError_t e = al::launch_sub(a, b, out, n, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "al_sub");
    return AL_OK;
}

int al_rmsnorm_bf16(void* x, const float* gamma, const float* bias, void* out, int64_t n_rows, int dim, float scale,
                    float eps, void* stream) {
    if (!x || !gamma || !out) return fail(AL_E_ARG, "al_rmsnorm_bf16: NULL argument");
    if (n_rows == 0) return AL_OK;
    if (n_rows < 0 || dim <= 0 || (dim & 7) != 0 || dim > 2048)
        return fail(AL_E_ARG, "al_rmsnorm_bf16: dim %d must be a multiple of 8, <= 2048", dim);
    if (((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(gamma) |
          reinterpret_cast<uintptr_t>(bias)) & 15) != 0)
        return fail(AL_E_ARG, "al_rmsnorm_bf16: pointers must be 16-byte aligned");
    cudaError_t e = al::launch_rmsnorm_bf16(x, gamma, bias, out, n_rows, dim, scale, eps, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "al_rmsnorm_bf16");
    return AL_OK;
}

int al_rotary_bf16(void* q, void* k, const float* cos_sin, int64_t n_rows, int heads, int dim_head, int64_t pos_div,
                   int pos_mod, void* stream) {
    if (!q || !k || !cos_sin) return fail(AL_E_ARG, "al_rotary_bf16: NULL argument");
    if (n_rows == 0) return AL_OK;
    if (n_rows < 0 || heads <= 0 || dim_head <= 0 || (dim_head & 7) != 0 || pos_div <= 0 || pos_mod <= 0)
        return fail(AL_E_ARG, "al_rotary_bf16: bad sizes (dim_head must be a multiple of 8)");
    if (((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(cos_sin)) & 15) != 0)
        return fail(AL_E_ARG, "al_rotary_bf16: pointers must be 16-byte aligned");
    cudaError_t e = al::launch_rotary_bf16(q, k, cos_sin, n_rows, heads, dim_head, pos_div, pos_mod, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "al_rotary_bf16");
    return AL_OK;
}

int al_gate_sigmoid_bf16(void* o, const void* gates, int64_t n_rows, int heads, int dim_head, void* stream) {
    if (!o || !gates) return fail(AL_E_ARG, "al_gate_sigmoid_bf16: NULL argument");
    if (n_rows == 0) return AL_OK;
    if (n_rows < 0 || heads <= 0 || dim_head <= 0 || (dim_head & 7) != 0)
        return fail(AL_E_ARG, "al_gate_sigmoid_bf16: bad sizes (dim_head must be a multiple of 8)");
    if ((reinterpret_cast<uintptr_t>(o) & 15) != 0) return fail(AL_E_ARG, "al_gate_sigmoid_bf16: o must be 16-byte aligned");
    cudaError_t e = al::launch_gate_bf16(o, gates, n_rows, heads, dim_head, heads, 0, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "al_gate_sigmoid_bf16");
    return AL_OK;
}

int al_gate_sigmoid_ld_bf16(void* o, const void* gates, int64_t gate_ld, int64_t n_rows, int heads, int dim_head, int fp16,
                            void* stream) {
    if (!o || !gates) return fail(AL_E_ARG, "al_gate_sigmoid_ld_bf16: NULL argument");
    if (n_rows == 0) return AL_OK;
    if (n_rows < 0 || heads <= 0 || dim_head <= 0 || (dim_head & 7) != 0 || gate_ld < heads || gate_ld > (1 << 20))
        return fail(AL_E_ARG, "al_gate_sigmoid_ld_bf16: bad sizes (dim_head must be a multiple of 8, gate_ld >= heads)");
    if ((reinterpret_cast<uintptr_t>(o) & 15) != 0) return fail(AL_E_ARG, "al_gate_sigmoid_ld_bf16: o must be 16-byte aligned");
    cudaError_t e = al::launch_gate_bf16(o, gates, n_rows, heads, dim_head, (int)gate_ld, fp16, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "al_gate_sigmoid_ld_bf16");
    return AL_OK;
}

int al_gelu_bf16(void* x, int64_t n, void* stream) {
    if (!x) return fail(AL_E_ARG, "al_gelu_bf16: NULL argument");
    if (n == 0) return AL_OK;
    if (n < 0 || (n & 7) != 0) return fail(AL_E_ARG, "al_gelu_bf16: n must be a non-negative multiple of 8");
    if ((reinterpret_cast<uintptr_t>(x) & 15) != 0) return fail(AL_E_ARG, "al_gelu_bf16: x must be 16-byte aligned");
    cudaError_t e = al::launch_gelu_bf16(x, n, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "al_gelu_bf16");
    return AL_OK;
}

int al_band_attention_bf16(const void* q, const void* k, const void* v, void* o, const void* gates, int64_t gate_ld,
                           const float* cos_sin, int64_t n_seq, int seq_len, int heads, int dim_head, float scale, int fp16,
                           void* stream) {
    if (!q || !k || !v || !o) return fail(AL_E_ARG, "al_band_attention_bf16: NULL argument");
    if (n_seq == 0) return AL_OK;
    if (n_seq < 0 || heads <= 0) return fail(AL_E_ARG, "al_band_attention_bf16: bad sizes");
    if (dim_head != 64 || seq_len < 1 || seq_len > 64)
        return fail(AL_E_UNSUPPORTED, "al_band_attention_bf16: dim_head %d / seq_len %d (needs 64 and 1..64)", dim_head, seq_len);
    if (((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
          reinterpret_cast<uintptr_t>(o)) & 15) != 0)
        return fail(AL_E_ARG, "al_band_attention_bf16: pointers must be 16-byte aligned");
    if (gates && gate_ld != 0 && (gate_ld < heads || gate_ld > (1 << 20)))
        return fail(AL_E_ARG, "al_band_attention_bf16: gate_ld %lld must be 0 (= heads) or >= heads", (long long)gate_ld);
    if (fp16 && cos_sin) return fail(AL_E_UNSUPPORTED, "al_band_attention_bf16: cos_sin (in-kernel rotary) is bfloat16-only");
    cudaError_t e = al::launch_band_attn_bf16(q, k, v, o, gates, cos_sin, n_seq, seq_len, heads, scale, (int)gate_ld, fp16,
                                              (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "al_band_attention_bf16");
    return AL_OK;
}

int al_time_attention_bf16(const void* q, const void* k, const void* v, void* o, const void* gates, int64_t gate_ld,
                           int64_t n_batch, int seq_len, int inner, int heads, int dim_head, float scale, int fp16,
                           void* stream) {
    if (!q || !k || !v || !o) return fail(AL_E_ARG, "al_time_attention_bf16: NULL argument");
    if (n_batch == 0) return AL_OK;
    if (n_batch < 0 || heads <= 0 || inner <= 0 || seq_len <= 0) return fail(AL_E_ARG, "al_time_attention_bf16: bad sizes");
    if (dim_head != 64) return fail(AL_E_UNSUPPORTED, "al_time_attention_bf16: dim_head %d (needs 64)", dim_head);
    if (gates && gate_ld != 0 && (gate_ld < heads || gate_ld > (1 << 20)))
        return fail(AL_E_ARG, "al_time_attention_bf16: gate_ld %lld must be 0 (= heads) or >= heads", (long long)gate_ld);
    cudaError_t ce = cudaSuccess;
    const char* msg = al::launch_time_attention(q, k, v, o, gates, gate_ld, n_batch, seq_len, inner, heads, dim_head, scale, fp16,
                                                (cudaStream_t)stream, &ce);
    if (!msg) return AL_OK;
    if (ce != cudaSuccess) return cuda_fail(ce, "al_time_attention_bf16");
    return fail(AL_E_ARG, "al_time_attention_bf16: %s", msg);
}

int al_gemm_bf16(const al_gemm_args* a, void* stream) {
    if (!a || !a->A || !a->W) return fail(AL_E_ARG, "al_gemm_bf16: NULL argument");
    if (a->M == 0) return AL_OK;
    if (a->epi != AL_GEMM_EPI_BF16 && a->epi != AL_GEMM_EPI_RESIDUAL && a->epi != AL_GEMM_EPI_GLU)
        return fail(AL_E_ARG, "al_gemm_bf16: bad epi %d", a->epi);
    cudaError_t ce = cudaSuccess;
    const char* msg = al::launch_gemm_bf16(*a, (cudaStream_t)stream, &ce);
    if (!msg) return AL_OK;
    if (ce != cudaSuccess) return cuda_fail(ce, "al_gemm_bf16");
    return fail(AL_E_ARG, "al_gemm_bf16: %s", msg);
}

int al_band_norm(const float* x, int64_t ldx, const float* gamma, const int32_t* band_off, int n_bands, void* out, int64_t ldo,
                 int64_t n_rows, float eps, int out_fp16, void* stream) {
    if (!x || !gamma || !band_off || !out) return fail(AL_E_ARG, "al_band_norm: NULL argument");
    if (n_rows == 0) return AL_OK;
    if (n_rows < 0 || n_bands <= 0 || n_bands > 1024 || ldx <= 0 || ldo <= 0) return fail(AL_E_ARG, "al_band_norm: bad sizes");
    cudaError_t e = al::launch_band_norm(x, ldx, gamma, band_off, n_bands, out, ldo, n_rows, eps, out_fp16, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "al_band_norm");
    return AL_OK;
}

int al_resid_prepare(const float* x_in, const float* bias, const float* gamma, float* x32, void* xb, float* ss,
                     int64_t n_rows, int dim, int ss_parts, float eps, int xb_fp16, void* stream) {
    if (!x_in || !x32 || !xb || !ss) return fail(AL_E_ARG, "al_resid_prepare: NULL argument");
    if (n_rows == 0) return AL_OK;
    if (n_rows < 0 || dim <= 0 || dim > 2048 || ss_parts <= 0 || ss_parts > 8 || dim % (8 * ss_parts) != 0)
        return fail(AL_E_ARG, "al_resid_prepare: dim %d must be a multiple of 8 * ss_parts (%d), <= 2048", dim, ss_parts);
    if (((reinterpret_cast<uintptr_t>(x_in) | reinterpret_cast<uintptr_t>(x32) | reinterpret_cast<uintptr_t>(xb) |
          reinterpret_cast<uintptr_t>(bias) | reinterpret_cast<uintptr_t>(gamma)) & 15) != 0)
        return fail(AL_E_ARG, "al_resid_prepare: pointers must be 16-byte aligned");
    cudaError_t e = al::launch_resid_prepare(x_in, bias, gamma, x32, xb, ss, n_rows, dim, ss_parts, eps, xb_fp16, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "al_resid_prepare");
    return AL_OK;
}

}  // extern "C"
