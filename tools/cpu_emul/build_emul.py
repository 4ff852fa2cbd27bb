"""Builds tools/cpu_emul/_emul.so: the [emul-begin]..[emul-end] regions of the index-logic kernels
(al_ola.cu, al_resample.cu) compiled for the host on top of cuda_emul.h, so their addressing, masking
and tiling can be checked on a box without a GPU (tests/test_kernel_emulation.py).  Test infrastructure."""
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "..", "..", "audiolab_b200", "csrc")
SO = os.path.join(HERE, "_emul.so")

GLUE = r'''
extern "C" void emul_ola_gather(const float* chunks, int n_chunks, int data_chunk0, int rows, int chunk_len,
                                const long long* offsets, const int* mult, const float* wtab, const int* tab_id,
                                long long n_total, long long p0, long long p1, const float* halo_in, int raw_out,
                                float eps, float scale, float* track, long long track_stride, int block) {
    const long long per_cta = (long long)block * kOlaVec;
    dim3 grid((unsigned)((p1 - p0 + per_cta - 1) / per_cta), (unsigned)std::min(rows, 2));
    emul_launch(grid, dim3(block), [&] {
        ola_gather_kernel(chunks, n_chunks, data_chunk0, rows, chunk_len, offsets, mult, wtab, tab_id, n_total, p0, p1,
                          halo_in, raw_out, eps, scale, track, track_stride);
    });
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
    srcs = [os.path.join(CSRC, f) for f in ("al_ola.cu", "al_resample.cu")]
    deps = srcs + [os.path.join(HERE, "cuda_emul.h"), __file__]
    if not force and os.path.exists(SO) and all(os.path.getmtime(SO) > os.path.getmtime(d) for d in deps):
        return SO
    gen = os.path.join(HERE, "_emul_gen.cpp")
    with open(gen, "w") as f:
        f.write('#include "cuda_emul.h"\n')
        for s in srcs:
            f.write(f"// ---- from {os.path.basename(s)}\n" + region(s) + "\n")
        f.write(GLUE)
    subprocess.run(["g++", "-std=c++20", "-O1", "-g", "-mfma", "-ffp-contract=fast", "-shared", "-fPIC", "-pthread",
                    "-I", HERE, "-o", SO, gen], check=True)
    return SO


if __name__ == "__main__":
    print(build(force=True))
