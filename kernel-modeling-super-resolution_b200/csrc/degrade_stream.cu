// degrade_stream.cu -- the generic TMA row-streaming kernel (degrade_stream_impl.cuh), replicate-padding instantiations,
// plus the shape test and the dispatcher of both padding modes.
#include "degrade_stream_impl.cuh"

namespace kmsr {

int launch_degrade_stream_zero(const DegradeArgs& a, cudaStream_t st);     // degrade_stream_zero.cu

namespace {
bool k_supported(int k) { return k == 11 || k == 13 || k == 15 || k == 21 || k == 31; }
}  // namespace

bool stream_shape_ok(const DegradeArgs& a, int down_mode, const char** why) {
    const Geometry& g = a.g;
    *why = "";
    if (down_mode != KMSR_DOWN_BOXMEAN) { *why = "box-mean downsampling only"; return false; }
    if (g.kh != g.kw || !k_supported(g.kh)) { *why = "square kernels of size 11, 13, 15, 21, 31"; return false; }
    const int S = g.stride;
    if (S != 2 && S != 4 && S != 8) { *why = "factor 2, 4 or 8"; return false; }
    if (a.patch_offsets) { *why = "scene windows not covered"; return false; }
    if (a.W != 64 && a.W != 128 && (a.W % 256 != 0 || a.W > 4096)) { *why = "W in {64, 128, 256*m}"; return false; }
    if (a.H < S || a.H % S != 0) { *why = "H a multiple of the factor"; return false; }
    if (a.H / S + (g.KW + S - 1) / S - 1 < 1) { *why = "empty output"; return false; }
    if (((uintptr_t)a.hr & 15) || (a.sH & 3) || (a.sC & 3) || (a.N > 1 && (a.sN & 3))) {
        *why = "HR base / strides not 16-byte aligned"; return false;
    }
    if (a.sH < a.W || a.sC < 1 || (a.N > 1 && a.sN < 1)) { *why = "non-positive strides"; return false; }
    if (a.N >= (1ll << 31) || a.N * a.C >= (1ll << 38)) { *why = "too many patches"; return false; }
    return true;
}

int launch_degrade_stream(const DegradeArgs& a, cudaStream_t st) {
    if (a.pad_mode != KMSR_PAD_REPLICATE) return launch_degrade_stream_zero(a, st);
    StreamArgs t;
    int sms = 0;
    const int rc = fill_stream_args(a, t, &sms);
    if (rc != KMSR_OK) return rc;
    set_algo("stream");
    return launch_stream_k<true>(a, t, a.g.kh, a.g.stride, sms, st);
}

}  // namespace kmsr
