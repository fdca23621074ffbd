// Host emulation of the device FFT passes: runs the 128 "threads" of each pass in a loop (forward
// or reverse order, to expose intra-pass hazards) over plain arrays standing in for shared memory.
#include <cstring>
#include <vector>
#include "../../spectrogram-midi_b200/csrc/fft2048.cuh"

using namespace aegis;

extern "C" int emul_fft2048(const float* in /*[2048][2]*/, const float* twiddle /*[2048][2]*/,
                            float* out /*[2048][2]*/, int reverse_order) {
    std::vector<cf> bufA(BUFA_SIZE), bufB(BUFB_SIZE), res(FFT_N);
    const cf* x = reinterpret_cast<const cf*>(in);
    const cf* tab = reinterpret_cast<const cf*>(twiddle);
    std::vector<FftTwiddles> tw(FFT_THREADS);
    for (int lt = 0; lt < FFT_THREADS; ++lt) fft2048_load_twiddles(lt, tab, tw[lt]);
    auto each = [&](auto&& body) {
        if (reverse_order) for (int lt = FFT_THREADS - 1; lt >= 0; --lt) body(lt);
        else for (int lt = 0; lt < FFT_THREADS; ++lt) body(lt);
    };
    each([&](int lt) {
        cf v[16];
        for (int a = 0; a < 16; ++a) v[a] = x[lt + 128 * a];
        fft2048_pass1(lt, v, tw[lt], bufA.data());
    });
    each([&](int lt) { fft2048_pass2(lt, tw[lt], bufA.data(), bufB.data()); });
    each([&](int lt) { fft2048_pass3(lt, bufB.data(), res.data()); });
    std::memcpy(out, res.data(), sizeof(cf) * FFT_N);
    return 0;
}

// in-place variant (one exchange buffer), with the load/store halves as separate "barrier" phases
extern "C" int emul_fft2048_inplace(const float* in, const float* twiddle, float* out, int reverse_order) {
    std::vector<cf> buf(BUFA_SIZE);
    std::vector<cf> regs(FFT_THREADS * 16);
    const cf* x = reinterpret_cast<const cf*>(in);
    const cf* tab = reinterpret_cast<const cf*>(twiddle);
    std::vector<FftTwiddles> tw(FFT_THREADS);
    for (int lt = 0; lt < FFT_THREADS; ++lt) fft2048_load_twiddles(lt, tab, tw[lt]);
    auto each = [&](auto&& body) {
        if (reverse_order) for (int lt = FFT_THREADS - 1; lt >= 0; --lt) body(lt);
        else for (int lt = 0; lt < FFT_THREADS; ++lt) body(lt);
    };
    each([&](int lt) {
        cf v[16];
        for (int a = 0; a < 16; ++a) v[a] = x[lt + 128 * a];
        fft2048_pass1(lt, v, tw[lt], buf.data());
    });
    each([&](int lt) { fft2048_pass2_load(lt, buf.data(), &regs[lt * 16]); });
    each([&](int lt) { fft2048_pass2_store(lt, &regs[lt * 16], tw[lt], buf.data()); });
    each([&](int lt) { fft2048_pass3_load(lt, buf.data(), &regs[lt * 16]); });
    each([&](int lt) { fft2048_pass3_store(lt, &regs[lt * 16], buf.data()); });
    std::memcpy(out, buf.data(), sizeof(cf) * FFT_N);
    return 0;
}
