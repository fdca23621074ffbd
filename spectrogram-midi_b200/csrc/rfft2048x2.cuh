// Real 2048-point FP32 FFT of TWO frames at once, one warp, packed f32x2 arithmetic.
//
// Blackwell issues FFMA2 / FADD2 / FMUL2 (two independent FP32 operations on a 64-bit register pair)
// at the lane rate of the scalar forms but from ONE issue slot.  Every arithmetic value here is a
// pair p2 = (frame A, frame B): the two frames never mix, so each frame's result is bit-identical to
// what it would be alone or with any other partner, and the whole transform costs half the issue
// slots of a scalar one.
//
// One frame: z[m] = x[2m] + i x[2m+1] (m < 1024), Z = FFT1024(z), then
//   X[k] = E[k] + W2048^k O[k],  E = (Z[k] + conj Z[1024-k]) / 2,  O = (Z[k] - conj Z[1024-k]) / 2i.
// FFT1024 = 32 x 32: lane j holds z[j + 32a] (a = register index), does a 32-point DFT over a in
// registers, multiplies by W1024^{jb}, exchanges through shared memory ONCE (within the warp:
// __syncwarp only), does the second 32-point DFT over j.  Index algebra:
//   n = j + 32a, k = b + 32c :  W1024^{nk} = W32^{ab} * W1024^{jb} * W32^{jc}
// 32-point DFT = 8 x radix-4 over n1, constant twiddles W32^{n2 k1}, 4 x radix-8 over n2
//   (n = 8 n1 + n2, k = k1 + 4 k2); output X[k1 + 4 k2] sits in register 8 k1 + pos8(k2).
//
// The lane functions are plain inline functions of (lane, pointers) so the same source compiles for
// the host (tests/emulation/rfft_emul.cpp runs the 32 lanes of each phase in a loop).
#pragma once

#if defined(__CUDACC__)
#define AEGIS_HD __host__ __device__ __forceinline__
#else
#define AEGIS_HD inline
#endif

namespace aegis {

// ---- packed pair of floats: x = frame A, y = frame B ------------------------------------------
struct alignas(8) p2 {
    float x, y;
};

#if defined(__CUDA_ARCH__)
AEGIS_HD p2 operator+(p2 a, p2 b) {
    const float2 r = __fadd2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y));
    return p2{r.x, r.y};
}
AEGIS_HD p2 operator-(p2 a, p2 b) {
    const float2 r = __fadd2_rn(make_float2(a.x, a.y), make_float2(-b.x, -b.y));
    return p2{r.x, r.y};
}
AEGIS_HD p2 operator*(p2 a, p2 b) {
    const float2 r = __fmul2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y));
    return p2{r.x, r.y};
}
AEGIS_HD p2 pfma(p2 a, p2 b, p2 c) {  // a * b + c
    const float2 r = __ffma2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y), make_float2(c.x, c.y));
    return p2{r.x, r.y};
}
#else
AEGIS_HD p2 operator+(p2 a, p2 b) { return p2{a.x + b.x, a.y + b.y}; }
AEGIS_HD p2 operator-(p2 a, p2 b) { return p2{a.x - b.x, a.y - b.y}; }
AEGIS_HD p2 operator*(p2 a, p2 b) { return p2{a.x * b.x, a.y * b.y}; }
AEGIS_HD p2 pfma(p2 a, p2 b, p2 c) { return p2{a.x * b.x + c.x, a.y * b.y + c.y}; }
#endif
AEGIS_HD p2 operator-(p2 a) { return p2{-a.x, -a.y}; }
AEGIS_HD p2 psplat(float s) { return p2{s, s}; }
AEGIS_HD p2 operator*(p2 a, float s) { return a * psplat(s); }
AEGIS_HD p2 pfma(p2 a, float s, p2 c) { return pfma(a, psplat(s), c); }

// complex value of both frames: 16 bytes, the unit of every shared-memory exchange
struct alignas(16) c2 {
    p2 re, im;
};
AEGIS_HD c2 operator+(c2 a, c2 b) { return c2{a.re + b.re, a.im + b.im}; }
AEGIS_HD c2 operator-(c2 a, c2 b) { return c2{a.re - b.re, a.im - b.im}; }

// forward 4-point DFT in place
AEGIS_HD void r4(c2& a0, c2& a1, c2& a2, c2& a3) {
    const c2 s02 = a0 + a2, d02 = a0 - a2, s13 = a1 + a3, d13 = a1 - a3;
    a0 = s02 + s13;
    a2 = s02 - s13;
    a1 = c2{d02.re + d13.im, d02.im - d13.re};  // d02 - i d13
    a3 = c2{d02.re - d13.im, d02.im + d13.re};  // d02 + i d13
}

constexpr float R_SQRT1_2 = 0.70710678118654752f;
// cos / sin of 2*pi*e/32
constexpr float R_C32[8] = {1.0f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f,
                            0.70710678118654752f, 0.55557023301960218f, 0.38268343236508977f, 0.19509032201612825f};

// v *= W32^E = exp(-2 pi i E / 32), E a compile-time constant in [0, 24)
template <int E>
AEGIS_HD void twiddle32(c2& v) {
    if constexpr (E == 0) {
    } else if constexpr (E == 8) {
        v = c2{v.im, -v.re};
    } else if constexpr (E == 16) {
        v = c2{-v.re, -v.im};
    } else if constexpr (E == 4) {
        v = c2{(v.re + v.im) * R_SQRT1_2, (v.im - v.re) * R_SQRT1_2};
    } else if constexpr (E == 12) {
        v = c2{(v.im - v.re) * R_SQRT1_2, -((v.re + v.im) * R_SQRT1_2)};
    } else if constexpr (E == 20) {
        v = c2{-((v.re + v.im) * R_SQRT1_2), (v.re - v.im) * R_SQRT1_2};
    } else {
        // cos(2 pi E/32), sin(2 pi E/32) from the first-octant table
        constexpr int q = E / 8, r = E % 8;
        constexpr float c0 = R_C32[r], s0 = (r == 0) ? 0.0f : R_C32[8 - r];
        constexpr float c = (q == 0) ? c0 : (q == 1) ? -s0 : (q == 2) ? -c0 : s0;
        constexpr float s = (q == 0) ? s0 : (q == 1) ? c0 : (q == 2) ? -s0 : -c0;
        // (re + i im)(c - i s) = (re c + im s) + i (im c - re s)
        const p2 re = pfma(v.im, s, v.re * c);
        const p2 im = pfma(v.re, -s, v.im * c);
        v = c2{re, im};
    }
}

// forward 8-point DFT in place; X[k] ends up at v[rpos8(k)]
AEGIS_HD constexpr int rpos8(int k) { return 2 * (k & 3) + (k >> 2); }

AEGIS_HD void r8(c2* v) {
    r4(v[0], v[2], v[4], v[6]);
    r4(v[1], v[3], v[5], v[7]);
    twiddle32<4>(v[3]);
    twiddle32<8>(v[5]);
    twiddle32<12>(v[7]);
#pragma unroll
    for (int k2 = 0; k2 < 4; ++k2) {
        const c2 s = v[2 * k2] + v[2 * k2 + 1];
        const c2 d = v[2 * k2] - v[2 * k2 + 1];
        v[2 * k2] = s;
        v[2 * k2 + 1] = d;
    }
}

// register that holds output k of fft32
AEGIS_HD constexpr int rpos32(int k) { return 8 * (k & 3) + rpos8(k >> 2); }

template <int K1, int N2>
AEGIS_HD void tw_row(c2* v) {
    if constexpr (N2 < 8) {
        twiddle32<K1 * N2>(v[N2 + 8 * K1]);
        tw_row<K1, N2 + 1>(v);
    }
}

// forward 32-point DFT in place; X[k] ends up at v[rpos32(k)]
AEGIS_HD void fft32(c2* v) {
#pragma unroll
    for (int n2 = 0; n2 < 8; ++n2) r4(v[n2], v[n2 + 8], v[n2 + 16], v[n2 + 24]);
    tw_row<1, 1>(v);
    tw_row<2, 1>(v);
    tw_row<3, 1>(v);
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) r8(v + 8 * k1);
}

// ---- shared-memory geometry of one warp's exchange area ----------------------------------------
constexpr int RF_N = 2048;             // real frame length
constexpr int RF_M = 1024;             // complex transform length
constexpr int RF_BINS = RF_N / 2 + 1;  // 1025
constexpr int RF_XPITCH = 33;          // exchange rows of 32 words (+1 pad): conflict-free 32-bit stores and loads
constexpr int RF_PLANE = 32 * RF_XPITCH;           // words per (frame, re|im) plane
// The exchange runs in two rounds (real parts, then imaginary parts) through [A plane][16 words][B plane]; frame B
// sits 16 banks after frame A so that the two half-warps of pass 2 (lanes 0-15 frame A, lanes 16-31 frame B,
// 16 consecutive words each) never collide.
constexpr int RF_FRAME_B = RF_PLANE + 16;
constexpr int RF_XCHG_WORDS = RF_FRAME_B + RF_PLANE;  // 2128 words

struct alignas(8) cf32 {
    float x, y;
};

// pass 1 = fft32 over a (v[a] holds z[lane + 32a] of both frames), then these twiddles:
// tw1[b * 32 + lane] = W1024^{lane * b}.  Leaves element (lane, b) of both frames in v[rpos32(b)].
AEGIS_HD void rfft_twiddle1(int lane, c2* v, const cf32* tw1) {
#pragma unroll
    for (int b = 1; b < 32; ++b) {
        const c2 y = v[rpos32(b)];
        const cf32 w = tw1[b * 32 + lane];
        const p2 wr = psplat(w.x), wi = psplat(w.y);
        v[rpos32(b)] = c2{pfma(y.re, wr, -(y.im * wi)), pfma(y.re, wi, y.im * wr)};
    }
}

// one exchange round, store side: the real (IM = false) or imaginary parts of element (lane, b) -> [lane][b]
template <bool IM>
AEGIS_HD void rfft_xstore(int lane, const c2* v, float* xbuf) {
    float* const row = xbuf + lane * RF_XPITCH;
#pragma unroll
    for (int b = 0; b < 32; ++b) {
        const p2 y = IM ? v[rpos32(b)].im : v[rpos32(b)].re;
        row[b] = y.x;
        row[RF_FRAME_B + b] = y.y;
    }
}

// pass 2 works on ONE frame per half-warp (h = lane >> 4) and packs two COLUMNS of that frame:
//   lo = column q, hi = column 32 - q   (q = lane & 15; q = 0: columns 0 and 16)
// so that after the 32-point DFT over j the conjugate partner of lo(V[c]) = Z[q + 32c], which is
// Z[1024 - q - 32c] = Z[(32 - q) + 32 (31 - c)], is hi(V[31 - c]) of the same thread: the real-spectrum
// split needs no further exchange.
AEGIS_HD constexpr int rfft_col1(int q) { return q ? 32 - q : 16; }

// one exchange round, load side: out[j] = (element (j, column q), element (j, column col1)) of this half-warp's frame
AEGIS_HD void rfft_xload(int lane, const float* xbuf, p2* out) {
    const int h = lane >> 4, q = lane & 15;
    const float* const p0 = xbuf + h * RF_FRAME_B + q;
    const float* const p1 = xbuf + h * RF_FRAME_B + rfft_col1(q);
#pragma unroll
    for (int j = 0; j < 32; ++j) out[j] = p2{p0[j * RF_XPITCH], p1[j * RF_XPITCH]};
}

// Split one conjugate pair (scalars of one frame): from Z[k] = (kr, ki), Z[1024-k] = (nr, ni) and W2048^k to the
// squared magnitudes of 2 X[k] and 2 X[1024-k] (the caller folds the 1/2 into the window).
AEGIS_HD void rfft_split_pair(float kr, float ki, float nr, float ni, cf32 w, float& pow_k, float& pow_n) {
    const float sr = kr + nr, si = ki - ni;  // S = Zk + conj Zn
    const float dr = kr - nr, di = ki + ni;  // D = Zk - conj Zn;  O' = D / i = (di, -dr)
    const float tr = di * w.x + dr * w.y;    // Re W O'
    const float ti = di * w.y - dr * w.x;    // Im W O'
    const float ar = sr + tr, ai = si + ti;  // 2 X[k]
    const float br = sr - tr, bi = si - ti;  // conj(2 X[1024-k])
    pow_k = ar * ar + ai * ai;
    pow_n = br * br + bi * bi;
}

// Twiddles of the packed split: for iteration c (0..15) and column pair q (0..15) the two lanes handle
//   lo: k = q + 32c                      hi: k = (32 - q) + 32c      (q = 0: lo k = 32c, hi k = 16 + 32c)
struct alignas(16) tw4 {
    float wr_lo, wr_hi, wi_lo, wi_hi;   // (cos, -sin)(2 pi k / 2048) of the lo / hi lane
};
AEGIS_HD constexpr int rfft_split_klo(int q, int c) { return q + 32 * c; }
AEGIS_HD constexpr int rfft_split_khi(int q, int c) { return (q ? 32 - q : 16) + 32 * c; }

// After fft32 of pass 2: emit |2X|^2 of every bin this thread owns through `emit(bin, power, is_k)`
// (is_k: the bin is congruent to q modulo 32, else to -q; the caller's swizzle offset depends only on that).
// Packed form: the lo lane splits the pair (Z[k], Z[1024 - k]) with k = q + 32c from lo(V[c]) and hi(V[31 - c]);
// the hi lane at the same time splits (Z[k'], Z[1024 - k']) with k' = (32 - q) + 32c from hi(V[c]) and
// lo(V[31 - c]): swapping the halves of V[31 - c] lines both pairs up lane by lane, so the sixteen operations of a
// split run as packed instructions on two bins at once.  The q == 0 lanes hold the self-conjugate columns 0 and 16:
// their partners are lo(V[32 - c]) and hi(V[31 - c]), picked with selects so the warp stays converged; k = 512 is
// the one bin left over.
template <class Emit>
AEGIS_HD void rfft_split_emit(int lane, const c2* v, const tw4* tw2p /*[16][16]*/, cf32 tw512, Emit&& emit) {
    const int q = lane & 15;
    const bool sp = (q == 0);
    const int klo = rfft_split_klo(q, 0), khi = rfft_split_khi(q, 0);
    const tw4* const twp = tw2p + q;
    constexpr int AHEAD = 4;   // twiddles are fetched AHEAD iterations before use (shared-memory latency)
    tw4 wq[AHEAD];
#pragma unroll
    for (int c = 0; c < AHEAD; ++c) wq[c] = twp[16 * c];
#pragma unroll
    for (int c = 0; c < 16; ++c) {
        const c2 A = v[rpos32(c)], B = v[rpos32(31 - c)], B0 = v[rpos32((32 - c) & 31)];
        const tw4 w = wq[c % AHEAD];
        if (c + AHEAD < 16) wq[c % AHEAD] = twp[16 * (c + AHEAD)];
        const p2 wre = p2{sp ? B0.re.x : B.re.y, sp ? B.re.y : B.re.x};   // Z[1024 - k] of the lo / hi lane
        const p2 wim = p2{sp ? B0.im.x : B.im.y, sp ? B.im.y : B.im.x};
        const p2 wr = p2{w.wr_lo, w.wr_hi}, wi = p2{w.wi_lo, w.wi_hi};
        const p2 sr = A.re + wre, si = A.im - wim;    // S = Zk + conj Zn
        const p2 dr = A.re - wre, di = A.im + wim;    // D = Zk - conj Zn;  O' = D / i = (di, -dr)
        const p2 tr = pfma(di, wr, dr * wi);          // Re W O'
        const p2 ti = pfma(di, wi, -(dr * wr));       // Im W O'
        const p2 ar = sr + tr, ai = si + ti;          // 2 X[k]
        const p2 br = sr - tr, bi = si - ti;          // conj(2 X[1024-k])
        const p2 pk = pfma(ar, ar, ai * ai), pn = pfma(br, br, bi * bi);
        emit(klo + 32 * c, pk.x, true);
        emit(RF_M - (klo + 32 * c), pn.x, false);
        emit(khi + 32 * c, pk.y, false);
        emit(RF_M - (khi + 32 * c), pn.y, true);
    }
    {
        const c2 a = v[rpos32(16)];
        float pk, pn;
        rfft_split_pair(a.re.x, a.im.x, a.re.x, a.im.x, tw512, pk, pn);
        if (sp) emit(RF_M / 2, pk, true);
    }
}

}  // namespace aegis
