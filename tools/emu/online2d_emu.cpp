// CPU emulation of the tiled OnlineStudy step (bayesloop_b200/csrc/online2d.cuh): the per-thread phases of
// online2d_phases.h are run for tid = 0 .. NT-1 with a barrier between the phases (exactly what the kernel does
// with __syncthreads()), tile by tile, and compared with a direct whole-grid computation: reflect correlation along
// axis 0, then axis 1 (scipy.ndimage.gaussian_filter1d semantics), clamp, x likelihood, sums.
// TEST INFRASTRUCTURE: g++ -O2 -std=c++17 tools/emu/online2d_emu.cpp -o /tmp/online2d_emu && /tmp/online2d_emu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../../bayesloop_b200/csrc/online2d_phases.h"

using namespace blg::o2;

static std::vector<double> weights(double sigma, int R, int M) {  // build_weights of common.cuh, zero padded
    std::vector<double> W(padded_taps(R, M), 0.0);
    if (R == 0) {
        W[0] = 1.0;
        return W;
    }
    double tot = 0.0;
    for (int j = 0; j <= 2 * R; ++j) {
        const double x = j - R;
        W[j] = std::exp(-0.5 / (sigma * sigma) * x * x);
        tot += W[j];
    }
    for (int j = 0; j <= 2 * R; ++j) W[j] /= tot;
    return W;
}

static int reflect_ref(long long i, int n) {  // independent restatement: walk the mirrored sequence
    while (i < 0 || i >= n) i = i < 0 ? -1 - i : 2LL * n - 1 - i;
    return (int)i;
}

struct Case {
    int n0, n1, R0, R1;
    double s0, s1;
    bool clamp;
    double limit;
    int mode;  // 0 convolution, 1 pointwise (state), 2 pointwise reset
};

template <int TH, int NT>
static int run(const Case &c, unsigned seed) {
    std::mt19937_64 rng(seed);
    std::uniform_real_distribution<double> U(0.0, 1.0);
    const int n0 = c.n0, n1 = c.n1, G = n0 * n1;
    std::vector<double> state(G), lik(G), base(G);
    double tot = 0.0;
    for (int g = 0; g < G; ++g) {
        state[g] = U(rng) * U(rng);
        tot += state[g];
        lik[g] = std::exp(-8.0 * U(rng));
        base[g] = U(rng) / G;
    }
    for (int g = 0; g < G; ++g) state[g] /= tot;
    const int R0 = c.mode == 0 ? c.R0 : 0, R1 = c.mode == 0 ? c.R1 : 0;
    const std::vector<double> W0 = weights(c.s0, R0, kM0), W1 = weights(c.s1, R1, kM1);

    // ---- direct reference
    std::vector<double> t0(G), t1(G), want(G);
    for (int i = 0; i < n0; ++i)
        for (int j = 0; j < n1; ++j) {
            double acc = 0.0;
            for (int k = -R0; k <= R0; ++k) acc = std::fma(W0[k + R0], state[(size_t)reflect_ref(i + k, n0) * n1 + j], acc);
            t0[(size_t)i * n1 + j] = acc;
        }
    for (int i = 0; i < n0; ++i)
        for (int j = 0; j < n1; ++j) {
            double acc = 0.0;
            for (int k = -R1; k <= R1; ++k) acc = std::fma(W1[k + R1], t0[(size_t)i * n1 + reflect_ref(j + k, n1)], acc);
            t1[(size_t)i * n1 + j] = acc;
        }
    double S1 = 0.0, S2 = 0.0;
    for (int g = 0; g < G; ++g) {
        double v = c.mode == 2 ? base[g] * 0.37 : (c.mode == 1 ? state[g] : t1[g]);
        if (c.clamp && c.mode != 2) v = v < c.limit ? c.limit : v;
        want[g] = v * lik[g];
        S1 += v;
        S2 += want[g];
    }

    // ---- emulation of the kernel, tile by tile
    const int r0max = R0 + 3, r1max = R1 + 2;  // the launch-wide maxima are larger than this hypothesis' radii
    const int P = (kTW + 2 * r1max) | 1, inRowsMax = TH + 2 * r0max;
    std::vector<double> in((size_t)inRowsMax * P), mid((size_t)TH * P), got(G, -1.0);
    const int tilesY = (n0 + TH - 1) / TH, tilesX = (n1 + kTW - 1) / kTW;
    auto likf = [&](int gi, int gj, long long g) {
        if (g != (long long)gi * n1 + gj) std::abort();
        return lik[g];
    };
    double s1 = 0.0, s2 = 0.0;
    for (int ty = 0; ty < tilesY; ++ty)
        for (int tx = 0; tx < tilesX; ++tx) {
            Tile<TH> t{n0, n1, ty * TH, tx * kTW, R0, R1, P};
            const double poison = std::nan("");
            for (auto &x : in) x = poison;  // anything the phases read without having written it shows up as NaN
            for (auto &x : mid) x = poison;
            const bool clamp = c.clamp && c.mode != 2;
            if (c.mode != 0) {
                for (int tid = 0; tid < NT; ++tid)
                    pointwise_phase(t, state.data(), c.mode == 2 ? base.data() : nullptr, 0.37, got.data(), clamp, c.limit,
                                    likf, tid, NT, s1, s2);
                continue;
            }
            for (int tid = 0; tid < NT; ++tid) load_phase(t, state.data(), in.data(), tid, NT);
            for (int tid = 0; tid < NT; ++tid) conv0_phase(t, in.data(), mid.data(), W0.data(), tid, NT);
            for (auto &x : in) x = poison;  // the last two phases must not touch the input buffer (the next tile loads into it)
            std::vector<double> outb((size_t)TH * kOutP, poison);
            for (int tid = 0; tid < NT; ++tid) conv1_phase(t, mid.data(), outb.data(), W1.data(), tid, NT);
            for (int tid = 0; tid < NT; ++tid)
                epilogue_phase(t, outb.data(), got.data(), clamp, c.limit, likf, tid, NT, s1, s2);
        }
    double worst = 0.0;
    for (int g = 0; g < G; ++g) {
        const double d = std::fabs(got[g] - want[g]) / (std::fabs(want[g]) + 1e-300);
        if (!(d <= worst)) worst = d;  // NaN-propagating maximum
    }
    const double e1 = std::fabs(s1 - S1) / S1, e2 = std::fabs(s2 - S2) / S2;
    const bool ok = worst < 1e-12 && e1 < 1e-12 && e2 < 1e-12;
    std::printf("%s tile %dx%d/%d threads grid %3dx%-3d R %2d/%-2d mode %d clamp %d: max rel err %.2e, sums %.1e %.1e\n",
                ok ? "ok  " : "FAIL", TH, kTW, NT, n0, n1, R0, R1, c.mode, (int)c.clamp, worst, e1, e2);
    return ok ? 0 : 1;
}

int main() {
    const Case cases[] = {
        {64, 64, 5, 7, 1.3, 1.9, false, 0, 0},     {150, 130, 31, 21, 7.7, 5.1, false, 0, 0},
        {150, 130, 0, 21, 0, 5.1, false, 0, 0},    {150, 130, 17, 0, 4.2, 0, false, 0, 0},
        {70, 200, 9, 33, 2.2, 8.3, true, 1e-7, 0}, {200, 75, 40, 3, 10.0, 0.8, false, 0, 0},
        {65, 129, 1, 1, 0.3, 0.3, false, 0, 0},    {40, 50, 45, 60, 11.0, 15.0, false, 0, 0},  // R >= n: multiple reflections
        {130, 150, 0, 0, 0, 0, true, 2e-5, 1},     {130, 150, 0, 0, 0, 0, false, 0, 1},
        {130, 150, 0, 0, 0, 0, true, 2e-5, 2},     {512, 512, 31, 21, 7.7, 5.1, true, 1e-9, 0},
    };
    int bad = 0;
    unsigned seed = 1;
    for (const Case &c : cases) bad += run<64, 512>(c, seed++);
    std::printf(bad ? "%d case(s) FAILED\n" : "all cases passed\n", bad);
    return bad ? 1 : 0;
}
