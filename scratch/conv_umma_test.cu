// conv_umma_test.cu -- stand-alone check + timing of csrc/selector_umma.cuh (tcgen05 convolutions of the selector) against
// an fp64 host convolution, one layer at a time on random data.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -lineinfo -o scratch/conv_umma_probe scratch/conv_umma_test.cu
//   scratch/conv_umma_probe [N patches = 64] [H = W = 256]
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../kernel-modeling-super-resolution_b200/csrc/selector_umma.cuh"

namespace kmsr {
std::atomic<long long> g_launches{0};
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vfprintf(stderr, fmt, ap);
    va_end(ap);
    fputc('\n', stderr);
}
}  // namespace kmsr
using namespace kmsr;

static float tf32_host(float v) {
    uint32_t b;
    memcpy(&b, &v, 4);
    b = (b + 0x1000u) & 0xFFFFE000u;
    float r;
    memcpy(&r, &b, 4);
    return r;
}
static uint32_t rng_state = 12345u;
static float urand() {
    rng_state = rng_state * 1664525u + 1013904223u;
    return (float)(rng_state >> 8) * (1.0f / 16777216.0f);
}

// [stage][4 chunks][2 COUT / 8][8][4]; rows nn < COUT: hi part of channel nn, nn >= COUT: lo part of channel nn - COUT
static std::vector<float> pack_weights(const std::vector<float>& w, int cin, int cout) {
    const int G = cin >= 16 ? cin / 16 : 1, S = cin == 5 ? 3 : 9 * G;
    std::vector<float> out((size_t)S * 4 * 2 * cout * 4, 0.0f);
    for (int s = 0; s < S; ++s)
        for (int c = 0; c < 4; ++c)
            for (int nn = 0; nn < 2 * cout; ++nn)
                for (int e = 0; e < 4; ++e) {
                    const int co = nn % cout;
                    float v = 0.0f;
                    if (cin == 5) {
                        const int kk = 4 * c + e;                 // k = 3 j + dx of triple T = 5 s + j (band = T / 3, dy = T % 3); 15: zero
                        if (kk < 15) { const int T = 5 * s + kk / 3; v = w[((size_t)co * cin + T / 3) * 9 + (T % 3) * 3 + kk % 3]; }
                    } else {
                        const int tap = s / G, g = s % G, ci = 16 * g + 4 * c + e;
                        v = w[((size_t)co * cin + ci) * 9 + tap];
                    }
                    const float hi = tf32_host(v), lo = tf32_host(v - hi);
                    out[(((size_t)s * 4 + c) * (2 * cout / 8) + nn / 8) * 32 + (nn % 8) * 4 + e] = nn < cout ? hi : lo;
                }
    return out;
}

template <int CIN, int COUT, bool POOL>
static int run_layer(int N, int H, int W, int sms) {
    const int Ho = H / 2, Wo = W / 2, tiles = Ho * Wo / 128;
    printf("layer %d -> %d%s: N = %d, %d x %d -> %d x %d, %d tiles per patch\n", CIN, COUT, POOL ? " (pooled)" : "", N, H, W, Ho, Wo, tiles);
    std::vector<float> w((size_t)COUT * CIN * 9), b(COUT);
    for (auto& v : w) v = (urand() - 0.5f) * 0.2f;
    for (auto& v : b) v = (urand() - 0.5f) * 0.1f;
    const std::vector<float> wst = pack_weights(w, CIN, COUT);
    const size_t in_elems = (size_t)N * CIN * H * W;
    std::vector<float> in(in_elems);
    for (auto& v : in) v = urand();
    const size_t out_elems = POOL ? (size_t)N * tiles * COUT : (size_t)N * Ho * Wo * COUT;
    float *d_in, *d_w, *d_b, *d_out;
    cudaMalloc(&d_in, in_elems * 4); cudaMalloc(&d_w, wst.size() * 4); cudaMalloc(&d_b, COUT * 4); cudaMalloc(&d_out, out_elems * 4);
    cudaMemcpy(d_in, in.data(), in_elems * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d_w, wst.data(), wst.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d_b, b.data(), COUT * 4, cudaMemcpyHostToDevice);
    cudaMemset(d_out, 0xff, out_elems * 4);
    umma::ConvUArgs a{};
    a.in = d_in; a.wst = d_w; a.bias = d_b; a.out = POOL ? nullptr : d_out; a.pool_part = POOL ? d_out : nullptr;
    a.H = H; a.W = W; a.Ho = Ho; a.Wo = Wo; a.tiles = tiles; a.passes = (long long)N * tiles / 2;
    int rc = umma::launch_conv_umma<CIN, COUT, POOL>(a, N, sms, 0);
    cudaError_t e = cudaDeviceSynchronize();
    if (rc != 0 || e != cudaSuccess) { printf("  launch rc %d, cuda: %s\n", rc, cudaGetErrorString(e)); return 1; }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        umma::launch_conv_umma<CIN, COUT, POOL>(a, N, sms, 0);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double flop = 2.0 * N * Ho * Wo * COUT * (CIN * 9.0);
    printf("  %.3f ms  (%.1f useful TFLOP/s, x3 = %.1f tensor TFLOP/s; %.2f ms per 4096 patches)\n", best, flop / best * 1e-9,
           3 * flop / best * 1e-9, best * 4096.0 / N);
    std::vector<float> out(out_elems);
    cudaMemcpy(out.data(), d_out, out_elems * 4, cudaMemcpyDeviceToHost);
    // reference: patches 0 and N - 1
    double worst = 0, scale = 0;
    long long bad = 0;
    for (int n : {0, N - 1}) {
        std::vector<double> pooled((size_t)tiles * COUT, 0.0);
        for (int oy = 0; oy < Ho; ++oy)
            for (int ox = 0; ox < Wo; ++ox)
                for (int co = 0; co < COUT; ++co) {
                    double s = b[co];
                    for (int ci = 0; ci < CIN; ++ci)
                        for (int dy = 0; dy < 3; ++dy)
                            for (int dx = 0; dx < 3; ++dx) {
                                const int iy = 2 * oy - 1 + dy, ix = 2 * ox - 1 + dx;
                                if (iy < 0 || ix < 0 || iy >= H || ix >= W) continue;
                                const float x = CIN == 5 ? in[(((size_t)n * CIN + ci) * H + iy) * W + ix]
                                                         : in[(((size_t)n * H + iy) * W + ix) * CIN + ci];
                                s += (double)x * (double)w[((size_t)co * CIN + ci) * 9 + dy * 3 + dx];
                            }
                    if (s < 0) s = 0;
                    const int P = oy * Wo + ox;
                    if (POOL) pooled[(size_t)(P / 128) * COUT + co] += s;
                    else {
                        const double got = out[(((size_t)n * Ho + oy) * Wo + ox) * COUT + co];
                        const double err = fabs(got - s);
                        if (!(err <= 1e-5 * (1.0 + fabs(s)))) { if (bad < 6) printf("  mismatch n %d (%d, %d) c %d: got %.8g want %.8g\n", n, oy, ox, co, got, s); ++bad; }
                        if (err > worst) worst = err;
                        if (fabs(s) > scale) scale = fabs(s);
                    }
                }
        if (POOL)
            for (int tl = 0; tl < tiles; ++tl)
                for (int co = 0; co < COUT; ++co) {
                    const double s = pooled[(size_t)tl * COUT + co], got = out[((size_t)n * tiles + tl) * COUT + co];
                    const double err = fabs(got - s);
                    if (!(err <= 1e-4 * (1.0 + fabs(s)))) { if (bad < 6) printf("  mismatch n %d tile %d c %d: got %.8g want %.8g\n", n, tl, co, got, s); ++bad; }
                    if (err > worst) worst = err;
                    if (fabs(s) > scale) scale = fabs(s);
                }
    }
    printf("  max abs error %.3g at scale %.3g, %lld over the bar\n", worst, scale, bad);
    cudaFree(d_in); cudaFree(d_w); cudaFree(d_b); cudaFree(d_out);
    return bad != 0;
}

int main(int argc, char** argv) {
    const int N = argc > 1 ? atoi(argv[1]) : 64, H = argc > 2 ? atoi(argv[2]) : 256;
    const int which = argc > 3 ? atoi(argv[3]) : 0;
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    int bad = 0;
    if (which == 0 || which == 2) bad |= run_layer<32, 64, false>(N, H / 2, H / 2, sms);
    if (which == 0 || which == 3) bad |= run_layer<64, 128, true>(N, H / 4, H / 4, sms);
    if (which == 0 || which == 1) bad |= run_layer<5, 32, false>(N, H, H, sms);
    printf(bad ? "FAILED\n" : "ok\n");
    return bad;
}
