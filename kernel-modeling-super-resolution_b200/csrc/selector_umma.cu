// selector_umma.cu -- SelectorNet inference (muti_kernel/train_gemini.py:14-39) through the tcgen05 convolution kernel of
// selector_umma.cuh: three launches of conv_umma_kernel (5 -> 32 -> 64 -> 128 channels, channel-last activations between
// them), then the pooling + linear kernel of selector.cu.  Patch shapes whose three output widths divide 128 (W = 64, 128,
// 256) with an even number of 128-pixel tiles per layer take this path; everything else stays on the mma.sync kernels.
#include "selector_umma.cuh"

namespace kmsr {

int launch_pool_fc(const float* part, int tiles, int C, float inv_area, const float* fc_w, const float* fc_b, int classes, long long N,
                   float* logits, cudaStream_t st);

namespace {
struct Dims { int h[4], w[4]; };
Dims dims_of(int H, int W) {
    Dims d;
    d.h[0] = H; d.w[0] = W;
    for (int l = 1; l < 4; ++l) { d.h[l] = d.h[l - 1] / 2; d.w[l] = d.w[l - 1] / 2; }
    return d;
}
inline long long pad256(long long b) { return (b + 255) / 256 * 256; }
}  // namespace

bool selector_umma_shape_ok(int H, int W) {
    if (H < 8 || W < 8 || H % 8 != 0 || W % 8 != 0) return false;
    const Dims d = dims_of(H, W);
    for (int l = 1; l < 4; ++l) {
        const int wo = d.w[l], ho = d.h[l];
        if (wo < 8 || wo > 128 || 128 % wo != 0) return false;            // an M tile is 128 / Wo whole output rows
        if ((long long)ho * wo % 256 != 0) return false;                     // two tiles per pass
    }
    return true;
}

// floats of one layer's weight stages: [stage][4 chunks][2 COUT / 8][8][4]
long long selector_umma_wfloats(int cin, int cout) {
    const int S = cin == 5 ? 3 : 9 * (cin / 16);
    return (long long)S * 4 * 2 * cout * 4;
}

long long selector_umma_workspace(long long N, int H, int W) {
    const Dims d = dims_of(H, W);
    long long bytes = 0;
    bytes += pad256(N * d.h[1] * d.w[1] * 32 * 4);
    bytes += pad256(N * d.h[2] * d.w[2] * 64 * 4);
    bytes += pad256(N * ((long long)d.h[3] * d.w[3] / 128) * 128 * 4);
    return bytes + 256;
}

int launch_selector_umma(const float* x, long long N, int H, int W, const float* w1, const float* b1, const float* w2, const float* b2,
                         const float* w3, const float* b3, const float* fc_w, const float* fc_b, float* logits, void* workspace,
                         long long workspace_bytes, cudaStream_t st) {
    if (N == 0) return KMSR_OK;
    KMSR_REQUIRE(selector_umma_shape_ok(H, W), KMSR_E_UNSUPPORTED,
                 "selector (tcgen05): %d x %d patches are not supported (W in {64, 128, 256}, even tile counts); use kmsr_selector_logits", H, W);
    KMSR_REQUIRE(workspace_bytes >= selector_umma_workspace(N, H, W), KMSR_E_INVALID, "selector (tcgen05): workspace of %lld bytes, %lld needed",
                 workspace_bytes, selector_umma_workspace(N, H, W));
    KMSR_REQUIRE(((uintptr_t)workspace & 255) == 0 && ((uintptr_t)w1 & 15) == 0 && ((uintptr_t)w2 & 15) == 0 && ((uintptr_t)w3 & 15) == 0 &&
                     ((uintptr_t)x & 15) == 0,
                 KMSR_E_ALIGN, "selector (tcgen05): workspace / weight blobs / input not aligned");
    KMSR_REQUIRE(N < (1ll << 31), KMSR_E_INVALID, "selector (tcgen05): too many patches");
    int dev = 0, sms = 0;
    KMSR_CUDA_OK(cudaGetDevice(&dev));
    KMSR_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const Dims d = dims_of(H, W);
    char* ws = reinterpret_cast<char*>(workspace);
    float* a1 = reinterpret_cast<float*>(ws);
    ws += pad256(N * d.h[1] * d.w[1] * 32 * 4);
    float* a2 = reinterpret_cast<float*>(ws);
    ws += pad256(N * d.h[2] * d.w[2] * 64 * 4);
    float* part = reinterpret_cast<float*>(ws);
    auto fill = [&](umma::ConvUArgs& c, int layer, const float* in, const float* w, const float* b, float* out, float* pp) {
        c.in = in; c.wst = w; c.bias = b; c.out = out; c.pool_part = pp;
        c.H = d.h[layer]; c.W = d.w[layer]; c.Ho = d.h[layer + 1]; c.Wo = d.w[layer + 1];
        c.tiles = c.Ho * c.Wo / 128;
        c.passes = N * c.tiles / umma::kTPP;
    };
    umma::ConvUArgs c{};
    fill(c, 0, x, w1, b1, a1, nullptr);
    int rc = umma::launch_conv_umma<5, 32, false>(c, N, sms, st);
    if (rc != KMSR_OK) return rc;
    fill(c, 1, a1, w2, b2, a2, nullptr);
    rc = umma::launch_conv_umma<32, 64, false>(c, N, sms, st);
    if (rc != KMSR_OK) return rc;
    fill(c, 2, a2, w3, b3, nullptr, part);
    rc = umma::launch_conv_umma<64, 128, true>(c, N, sms, st);
    if (rc != KMSR_OK) return rc;
    return launch_pool_fc(part, c.tiles, 128, 1.0f / (float)(d.h[3] * d.w[3]), fc_w, fc_b, 10, N, logits, st);
}

}  // namespace kmsr
