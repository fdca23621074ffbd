// Host emulation of the warp-level packed real FFT (csrc/rfft2048x2.cuh): the 32 "lanes" of each
// phase run in a loop (forward or reverse order) over plain arrays standing in for shared memory.
#include <cmath>
#include <cstring>
#include <vector>
#include "../../spectrogram-midi_b200/csrc/rfft2048x2.cuh"

using namespace aegis;

// frames: [2][2048] real (already windowed); twiddle: [2048][2] (cos, -sin)(2 pi k / 2048)
// out: [2][1025] magnitudes of the two frames' spectra (x 2, as the kernel computes before its 1/2 fold)
extern "C" int emul_rfft2048x2(const float* frames, const float* twiddle, float* out, int reverse_order) {
    const cf32* tab = reinterpret_cast<const cf32*>(twiddle);
    std::vector<cf32> tw1(32 * 32);
    std::vector<tw4> tw2p(16 * 16);
    for (int b = 0; b < 32; ++b)
        for (int j = 0; j < 32; ++j) tw1[b * 32 + j] = tab[(2 * j * b) & 2047];
    for (int c = 0; c < 16; ++c)
        for (int q = 0; q < 16; ++q) {
            const cf32 lo = tab[rfft_split_klo(q, c)], hi = tab[rfft_split_khi(q, c)];
            tw2p[c * 16 + q] = tw4{lo.x, hi.x, lo.y, hi.y};
        }
    std::vector<float> buf(RF_XCHG_WORDS);
    std::vector<p2> nre(32 * 32);
    std::vector<c2> regs(32 * 32);
    auto each = [&](auto&& body) {
        if (reverse_order) for (int l = 31; l >= 0; --l) body(l);
        else for (int l = 0; l < 32; ++l) body(l);
    };
    const float* fa = frames;
    const float* fb = frames + 2048;
    each([&](int lane) {
        c2* v = &regs[lane * 32];
        for (int a = 0; a < 32; ++a) {
            const int m = lane + 32 * a;
            v[a] = c2{p2{fa[2 * m], fb[2 * m]}, p2{fa[2 * m + 1], fb[2 * m + 1]}};
        }
        fft32(v);
        rfft_twiddle1(lane, v, tw1.data());
        rfft_xstore<false>(lane, v, buf.data());
    });
    each([&](int lane) { rfft_xload(lane, buf.data(), &nre[lane * 32]); });
    each([&](int lane) { rfft_xstore<true>(lane, &regs[lane * 32], buf.data()); });
    each([&](int lane) {
        c2* v = &regs[lane * 32];
        p2 nim[32];
        rfft_xload(lane, buf.data(), nim);
        for (int j = 0; j < 32; ++j) v[j] = c2{nre[lane * 32 + j], nim[j]};
    });
    std::vector<int> written(2 * 1025, 0);
    each([&](int lane) {
        c2* v = &regs[lane * 32];
        fft32(v);
        const int h = lane >> 4;
        rfft_split_emit(lane, v, tw2p.data(), tab[512], [&](int k, float pw, bool) {
            out[h * 1025 + k] = std::sqrt(pw);
            written[h * 1025 + k] += 1;
        });
    });
    for (int i = 0; i < 2 * 1025; ++i)
        if (written[i] != 1) return 1 + i;  // every bin of both frames exactly once
    return 0;
}
