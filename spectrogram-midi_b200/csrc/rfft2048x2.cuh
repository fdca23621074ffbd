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

// ---- shared-memory geometry of one warp's exchange buffer --------------------------------------
constexpr int RF_N = 2048;             // real frame length
constexpr int RF_M = 1024;             // complex transform length
constexpr int RF_BINS = RF_N / 2 + 1;  // 1025
constexpr int RF_XPITCH = 33;          // exchange rows of 32 c2 (+1 pad): conflict-free 128-bit stores and loads
constexpr int RF_WARP_BUF = 32 * RF_XPITCH;  // c2 elements (16896 B)

struct alignas(8) cf32 {
    float x, y;
};

// pass 1: v[a] holds z[lane + 32a] of both frames.  tw1[b * 32 + lane] = W1024^{lane * b}.
AEGIS_HD void rfft_pass1(int lane, c2* v, const cf32* tw1, c2* xbuf) {
    fft32(v);
#pragma unroll
    for (int b = 0; b < 32; ++b) {
        c2 y = v[rpos32(b)];
        if (b) {
            const cf32 w = tw1[b * 32 + lane];
            const p2 wr = psplat(w.x), wi = psplat(w.y);
            const p2 re = pfma(y.re, wr, -(y.im * wi));
            const p2 im = pfma(y.re, wi, y.im * wr);
            y = c2{re, im};
        }
        xbuf[lane * RF_XPITCH + b] = y;
    }
}

AEGIS_HD void rfft_pass2_load(int lane, const c2* xbuf, c2* v) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = xbuf[j * RF_XPITCH + lane];
}

// pass 2 (after every lane finished rfft_pass2_load): Z[lane + 32c] -> zbuf (natural order, aliases xbuf)
AEGIS_HD void rfft_pass2_store(int lane, c2* v, c2* zbuf) {
    fft32(v);
#pragma unroll
    for (int c = 0; c < 32; ++c) zbuf[lane + 32 * c] = v[rpos32(c)];
}

// Split one conjugate pair: from Z[k], Z[1024-k] and W2048^k to 2 X[k], 2 X[1024-k]; returns the two
// squared magnitudes (of 2X; the caller folds the 1/2 into the window).
AEGIS_HD void rfft_split_pair(c2 zk, c2 zn, cf32 w, p2& pow_k, p2& pow_n) {
    const p2 sr = zk.re + zn.re, si = zk.im - zn.im;  // S = Zk + conj Zn
    const p2 dr = zk.re - zn.re, di = zk.im + zn.im;  // D = Zk - conj Zn;  O' = D / i = (di, -dr)
    const p2 wr = psplat(w.x), wi = psplat(w.y);
    const p2 tr = pfma(di, wr, dr * wi);              // Re W O' = wr di + wi dr
    const p2 ti = pfma(di, wi, -(dr * wr));           // Im W O' = wi di - wr dr
    const p2 ar = sr + tr, ai = si + ti;              // 2 X[k]
    const p2 br = sr - tr, bi = si - ti;              // conj(2 X[1024-k])
    pow_k = pfma(ar, ar, ai * ai);
    pow_n = pfma(br, br, bi * bi);
}

}  // namespace aegis
