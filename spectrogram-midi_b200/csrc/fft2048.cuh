// 2048-point complex FP32 FFT, 128 threads x 16 points, three register-radix passes (16 x 16 x 8)
// with two shared-memory exchanges.  Shared by the STFT kernel (two real frames packed into one
// complex transform) and the YIN kernel (forward transforms + one packed inverse).
//
// The per-thread pass bodies are plain inline functions of (thread index, shared buffers) so the
// same source compiles for the device and -- with AEGIS_HOST_EMULATION -- for a host harness that
// runs the "threads" of each pass in a loop (tests/test_fft_emulation.py).  No warp intrinsics here.
//
// Index algebra (W = exp(-2*pi*i/2048)):
//   n = t + 128a, k = b + 16c      : W^{nk} = W^{tb} * W128^{tc} * W16^{ab}
//   t = t0 + 8t1, c = d + 16e      : W128^{tc} = W128^{t0 d} * W8^{t0 e} * W16^{t1 d}
//   pass 1: thread t      : 16-pt DFT over a, times W^{tb}          -> bufA[b][t]
//   pass 2: thread (b,t0) : 16-pt DFT over t1, times W128^{t0 d}    -> bufB[t0][d][b]
//   pass 3: item (b,d)    : 8-pt DFT over t0                        -> out[b + 16d + 256e]
// Shared layouts are padded so every LDS.64/STS.64 half-warp hits 16 distinct bank pairs.
#pragma once

#if defined(__CUDACC__)
#define AEGIS_HD __host__ __device__ __forceinline__
#else
#define AEGIS_HD inline
#endif

namespace aegis {

struct alignas(8) cf {
    float x, y;
};

constexpr int FFT_N = 2048;
constexpr int FFT_THREADS = 128;
constexpr int BUFA_PITCH = 136;               // b-major rows of 128 (+8 pad)
constexpr int BUFA_SIZE = 16 * BUFA_PITCH;    // 2176 cf
constexpr int BUFB_PITCH = 258;               // t0-major rows of 256 (+2 pad)
constexpr int BUFB_SIZE = 8 * BUFB_PITCH;     // 2064 cf

AEGIS_HD cf cadd(cf a, cf b) { return cf{a.x + b.x, a.y + b.y}; }
AEGIS_HD cf csub(cf a, cf b) { return cf{a.x - b.x, a.y - b.y}; }
AEGIS_HD cf cmul(cf a, cf w) { return cf{a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x}; }

// forward 4-point DFT in place: (a0,a1,a2,a3) -> (X0,X1,X2,X3)
AEGIS_HD void bfly4(cf& a0, cf& a1, cf& a2, cf& a3) {
    cf s02 = cadd(a0, a2), d02 = csub(a0, a2), s13 = cadd(a1, a3), d13 = csub(a1, a3);
    a0 = cadd(s02, s13);
    a2 = csub(s02, s13);
    a1 = cf{d02.x + d13.y, d02.y - d13.x};  // d02 - i*d13
    a3 = cf{d02.x - d13.y, d02.y + d13.x};  // d02 + i*d13
}

constexpr float C_PI8 = 0.92387953251128674f;   // cos(pi/8)
constexpr float S_PI8 = 0.38268343236508977f;   // sin(pi/8)
constexpr float RSQRT2 = 0.70710678118654752f;

// forward 16-point DFT in place; X[k] ends up at v[pos16(k)]
AEGIS_HD constexpr int pos16(int k) { return 4 * (k & 3) + (k >> 2); }

AEGIS_HD void fft16(cf* v) {
#pragma unroll
    for (int n1 = 0; n1 < 4; ++n1) bfly4(v[n1], v[n1 + 4], v[n1 + 8], v[n1 + 12]);
    // v[n1 + 4*k2] *= W16^{n1*k2}
    v[5] = cmul(v[5], cf{C_PI8, -S_PI8});                                       // W^1
    v[9] = cf{(v[9].x + v[9].y) * RSQRT2, (v[9].y - v[9].x) * RSQRT2};          // W^2
    v[13] = cmul(v[13], cf{S_PI8, -C_PI8});                                     // W^3
    v[6] = cf{(v[6].x + v[6].y) * RSQRT2, (v[6].y - v[6].x) * RSQRT2};          // W^2
    v[10] = cf{v[10].y, -v[10].x};                                              // W^4 = -i
    v[14] = cf{(v[14].y - v[14].x) * RSQRT2, -(v[14].x + v[14].y) * RSQRT2};    // W^6
    v[7] = cmul(v[7], cf{S_PI8, -C_PI8});                                       // W^3
    v[11] = cf{(v[11].y - v[11].x) * RSQRT2, -(v[11].x + v[11].y) * RSQRT2};    // W^6
    v[15] = cmul(v[15], cf{-C_PI8, S_PI8});                                     // W^9
#pragma unroll
    for (int k2 = 0; k2 < 4; ++k2) bfly4(v[4 * k2], v[4 * k2 + 1], v[4 * k2 + 2], v[4 * k2 + 3]);
}

// forward 8-point DFT in place; X[k] ends up at v[pos8(k)]
AEGIS_HD constexpr int pos8(int k) { return 2 * (k & 3) + (k >> 2); }

AEGIS_HD void fft8(cf* v) {
    bfly4(v[0], v[2], v[4], v[6]);
    bfly4(v[1], v[3], v[5], v[7]);
    // v[1 + 2*k2] *= W8^{k2}
    v[3] = cf{(v[3].x + v[3].y) * RSQRT2, (v[3].y - v[3].x) * RSQRT2};
    v[5] = cf{v[5].y, -v[5].x};
    v[7] = cf{(v[7].y - v[7].x) * RSQRT2, -(v[7].x + v[7].y) * RSQRT2};
#pragma unroll
    for (int k2 = 0; k2 < 4; ++k2) {
        cf s = cadd(v[2 * k2], v[2 * k2 + 1]);
        cf d = csub(v[2 * k2], v[2 * k2 + 1]);
        v[2 * k2] = s;
        v[2 * k2 + 1] = d;
    }
}

// Per-thread twiddles, loop invariant across transforms: tw1[b] = W^{t*b}, tw2[d] = W128^{t0*d}.
struct FftTwiddles {
    cf tw1[16];
    cf tw2[16];
};

AEGIS_HD void fft2048_load_twiddles(int lt, const cf* table /*[2048] (cos,-sin)*/, FftTwiddles& tw) {
    const int t0 = lt & 7;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        tw.tw1[j] = table[(lt * j) & (FFT_N - 1)];
        tw.tw2[j] = table[(16 * t0 * j) & (FFT_N - 1)];
    }
}

// pass 1: v[a] must hold x[lt + 128a] on entry
AEGIS_HD void fft2048_pass1(int lt, cf* v, const FftTwiddles& tw, cf* bufA) {
    fft16(v);
#pragma unroll
    for (int b = 0; b < 16; ++b) {
        cf y = v[pos16(b)];
        if (b) y = cmul(y, tw.tw1[b]);
        bufA[b * BUFA_PITCH + lt] = y;
    }
}

AEGIS_HD void fft2048_pass2(int lt, const FftTwiddles& tw, const cf* bufA, cf* bufB) {
    const int bq = lt >> 3, t0 = lt & 7;
    cf v[16];
#pragma unroll
    for (int t1 = 0; t1 < 16; ++t1) v[t1] = bufA[bq * BUFA_PITCH + t0 + 8 * t1];
    fft16(v);
#pragma unroll
    for (int d = 0; d < 16; ++d) {
        cf u = v[pos16(d)];
        if (d) u = cmul(u, tw.tw2[d]);
        bufB[t0 * BUFB_PITCH + d * 16 + bq] = u;
    }
}

// pass 3: natural-order spectrum out[0..2047]
AEGIS_HD void fft2048_pass3(int lt, const cf* bufB, cf* out) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int item = lt + 128 * r;  // = b + 16*d
        cf v[8];
#pragma unroll
        for (int t0 = 0; t0 < 8; ++t0) v[t0] = bufB[t0 * BUFB_PITCH + item];
        fft8(v);
#pragma unroll
        for (int e = 0; e < 8; ++e) out[item + 256 * e] = v[pos8(e)];
    }
}

// ---- in-place variants: one exchange buffer (BUFA_SIZE) per transform.  The caller separates
// "load" and "store" halves of passes 2 and 3 with a barrier so every read of the previous layout
// has completed before the buffer is overwritten with the next one.
AEGIS_HD void fft2048_pass2_load(int lt, const cf* buf, cf* v) {
    const int bq = lt >> 3, t0 = lt & 7;
#pragma unroll
    for (int t1 = 0; t1 < 16; ++t1) v[t1] = buf[bq * BUFA_PITCH + t0 + 8 * t1];
}

AEGIS_HD void fft2048_pass2_store(int lt, cf* v, const FftTwiddles& tw, cf* buf) {
    const int bq = lt >> 3, t0 = lt & 7;
    fft16(v);
#pragma unroll
    for (int d = 0; d < 16; ++d) {
        cf u = v[pos16(d)];
        if (d) u = cmul(u, tw.tw2[d]);
        buf[t0 * BUFB_PITCH + d * 16 + bq] = u;
    }
}

AEGIS_HD void fft2048_pass3_load(int lt, const cf* buf, cf* v /*[16]: two 8-point problems*/) {
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int t0 = 0; t0 < 8; ++t0) v[8 * r + t0] = buf[t0 * BUFB_PITCH + lt + 128 * r];
}

AEGIS_HD void fft2048_pass3_store(int lt, cf* v, cf* out) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        fft8(v + 8 * r);
#pragma unroll
        for (int e = 0; e < 8; ++e) out[lt + 128 * r + 256 * e] = v[8 * r + pos8(e)];
    }
}

static_assert(BUFB_SIZE <= BUFA_SIZE && FFT_N <= BUFA_SIZE, "in-place exchange needs one BUFA_SIZE buffer");

}  // namespace aegis
