// degrade_stream_zero.cu -- zero-padding instantiations (train_gemini.py:128) of the generic TMA row-streaming kernel.
#include "degrade_stream_impl.cuh"

namespace kmsr {

int launch_degrade_stream_zero(const DegradeArgs& a, cudaStream_t st) {
    StreamArgs t;
    int sms = 0;
    const int rc = fill_stream_args(a, t, &sms);
    if (rc != KMSR_OK) return rc;
    set_algo("stream");
    return launch_stream_k<false>(a, t, a.g.kh, a.g.stride, sms, st);
}

}  // namespace kmsr
