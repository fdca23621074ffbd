// K3: pitch/voicing HMM Viterbi decode for pYIN (librosa.sequence.viterbi inside librosa.pyin;
// call sites aegis_engine.py:63,67,190,216, aegis_engine_core/worker.py:9-15,
// aegis_engine_financial.py:63-69; algorithm SURVEY.md Appendix A.5 steps 10-12).
//
// States: 2*n_bins (voiced bins 0..n-1, unvoiced n..2n-1).  librosa evaluates a dense
// [2n x 2n] max-plus product per frame in float64 with log(p + tiny) everywhere, so "impossible"
// transitions/observations cost a finite log(tiny) and still compete.  This kernel is exact with
// respect to that definition (same float64 adds, first-index argmax) but never touches the dense
// matrix.  One CTA per clip, one thread per pitch bin (it owns the voiced and the unvoiced state of
// that bin), sequential over frames with one barrier per frame; V is ping-ponged in shared memory.
//
// Round 2 design: every decision that steers control flow is WARP-UNIFORM (a lane = one destination
// bin, the 32 destinations of a warp walk the same source chunks together), so there is no divergence
// and the edge warps do the same work as the interior ones:
//   1. in-band sources (|b - b'| <= half_width) are visited in aligned chunks of 8 source bins.
//      A lane-parallel PREFILTER (lane r tests chunk r of the warp's window with a bound that holds
//      for all 32 destinations) leaves a few chunks; for those a per-lane exact bound decides, and a
//      chunk is evaluated (8 float64 adds + leftmost-max tournament per lane) if any lane needs it;
//   2. out-of-band sources all carry log(tiny): only a voicing block's leftmost global maximum can
//      matter, and only when max(V) + log(tiny) can reach the warp's weakest lower bound -- checked
//      with one REDUX on the high words; the exact leftmost argmax is computed on demand;
//   3. DOMINANCE: a voiced state without a candidate (observation log(tiny)) in a frame whose
//      unvoiced observation is > log(tiny) + 10 (i.e. voiced_prob < 1) is strictly dominated by its
//      unvoiced twin, as a source of every transition of the next frame and as the final state:
//        V[0][b] <= V[1][b] - (lp_unvoiced - log tiny) + log(99) + ulps,  and the two states'
//        transitions to any destination differ by at most log(99) (or are both log(tiny)),
//      so its candidate is smaller than the twin's by > 0.8: it never wins, never ties.  Such states
//      are not computed at all (V = -inf, no back-pointer): in a frame with voiced_prob < 1 only the
//      warps that hold a candidate bin evaluate voiced destinations.
// Rows of the transition table that differ in the last ulp (librosa's pairwise row sums) are kept as
// "variants": the few interior ones in shared memory (padded with -inf so that no lane needs a range
// check), the truncated edge rows in global memory (a source row is warp-uniform inside a chunk, the
// 32 lanes read consecutive offsets: coalesced, L1 resident).
// Back-pointers stream to global (uint16); a second kernel walks them, one thread per clip.
//
// Precondition: observation probabilities lie in [0, 1] (every V is then negative, which the
// high-word bounds rely on); K2 guarantees it.
#include <cfloat>
#include <cmath>
#include "common.cuh"

namespace aegis {

constexpr int VT_MAX_BINS = 512;
constexpr int VT_HALO = 64;            // V halo: sources outside [0, n) read -inf; needs half_width + 7 <= 64
constexpr int VT_MAX_HW = 50;
constexpr int VT_MAX_W = 2 * VT_MAX_HW + 1;
constexpr int VT_SMEM_VARIANTS = 6;
constexpr int VT_MAX_WARPS = VT_MAX_BINS / 32;
constexpr int VT_CHUNK = 8;            // sources are pruned in aligned chunks of 8 bins
constexpr int VT_CHUNK_PAD = 8;        // chunk indices -8 .. (512/8 + 8)
constexpr int VT_N_CHUNKS = VT_MAX_BINS / VT_CHUNK + 2 * VT_CHUNK_PAD;
constexpr int VT_LT_PAD = 40;          // table offsets reach -38 .. W + 37 for lanes outside a chunk's band
constexpr int VT_LT_PITCH = VT_MAX_W + 2 * VT_LT_PAD + 1;   // odd number of doubles
constexpr int VT_UBR_PAD = 32;         // chunk offsets reach -31 .. W + 37
constexpr int VT_UBR_SIZE = VT_MAX_W + VT_UBR_PAD + 40;

struct VitSmem {
    double V[2][2][VT_MAX_BINS + 2 * VT_HALO];     // [ping][voicing][halo | bins | halo], halo = -inf
    double M[2][2][VT_N_CHUNKS];                   // max of V over each aligned chunk of 8 bins + cdub (-inf outside)
    double obs_lp[2][VT_MAX_BINS];                 // log(obs + tiny) of the voiced states
    double lt[VT_SMEM_VARIANTS][2][VT_LT_PITCH];   // interior transition variants [variant][same|switch][pad + offset], -inf padding
    double ubr[2][VT_UBR_SIZE];                    // [same|switch][pad + q]: max over the interior variants and over offsets q-7..q
    double ubrmax[2][32];                          // [same|switch][r]: max over the 32 destinations of a warp of ubr for window chunk r
    double cdub[VT_N_CHUNKS];                      // per chunk: how far its (edge) rows' tables exceed the interior ones
    unsigned seg_hi[2][2][VT_MAX_WARPS];           // per warp: smallest high word of V (= upper bound of the warp's maximum)
    unsigned char rowvar[VT_MAX_BINS + 2 * VT_HALO];
    double lp_u[2];                                // log-observation of the unvoiced states of frame t (same for all bins)
};

struct VA {
    double v;
    int a;
};

// leftmost-max combine under a butterfly exchange: every lane ends with (max value, lowest index attaining it)
__device__ __forceinline__ VA butterfly_leftmost(VA x, int mask) {
    const double ov = __shfl_xor_sync(0xffffffffu, x.v, mask);
    const int oa = __shfl_xor_sync(0xffffffffu, x.a, mask);
    if (ov > x.v || (ov == x.v && oa < x.a)) { x.v = ov; x.a = oa; }
    return x;
}

// (a, ia) has the lower source index: (b, ib) wins only when strictly greater -- numpy's first-index argmax
__device__ __forceinline__ void take_later(double& a, int& ia, double b, int ib) {
    if (b > a) { a = b; ia = ib; }
}

// Bounds from high words.  All arguments are negative doubles or -inf, for which the bit pattern grows with the magnitude.
// <= the smallest of the lanes' values whose high words went into `hmax` (= the largest high word)
__device__ __forceinline__ double lower_bound_from_hi(unsigned hmax) {
    return hmax >= 0xFFF00000u ? -INFINITY : __hiloint2double(static_cast<int>(hmax), -1);
}
// >= the largest of the values whose high words went into `hmin` (= the smallest high word); -inf stays -inf
__device__ __forceinline__ double upper_bound_from_hi(unsigned hmin) {
    return __hiloint2double(static_cast<int>(hmin), 0);
}

// exact leftmost maximum of one voicing block of V (on demand: only when an out-of-band source could matter)
__device__ __noinline__ VA block_leftmost_max(const double* __restrict__ Vc, int n, int lane) {
    VA x{-INFINITY, 0x7fffffff};
    for (int i = lane; i < n; i += 32) {
        const double v = Vc[i];
        if (v > x.v) { x.v = v; x.a = i; }
    }
#pragma unroll
    for (int m = 1; m < 32; m <<= 1) x = butterfly_leftmost(x, m);
    return x;
}

// One frame of the recursion for the 32 destinations of a warp.  BOTH: voiced destinations too (best0 / arg0).
// Exactness of the pruning (everything is float64, first-index argmax as in numpy):
//  * a chunk of 8 sources is skipped for a destination when  chunk_max + max(lt over the chunk's offsets, all interior
//    variants) [+ the excess of edge rows, cdub]  is < L, the value of a real candidate of that destination (its own bin
//    or the source that won in the previous frame), or <= the running best of lower-index sources: floating-point
//    addition is monotone, so no source of the chunk can reach the maximum or win a tie.  The prefilter uses the same
//    bound maximised over the warp's 32 destinations against a lower bound of the smallest L of the warp.
//  * chunks that survive are evaluated exactly, in ascending source order (voiced block first), ties to the lower index.
//  * out-of-band sources of a block: let g be the block's leftmost maximum.  If g lies inside the band its in-band
//    candidate (>= V[g] - 14) beats every out-of-band one (<= V[g] - 708); otherwise g is the best out-of-band source
//    of its side and the other side can at best tie with a higher index.
template <int HW, bool BOTH>
__device__ __forceinline__ void viterbi_frame_step(const VitSmem& s, const aegis_viterbi_params& p, const int cur, const int d,
                                                   const int lane, const int warp, const bool live, const int n, const int nsv,
                                                   const int n_warps, const int prev0, const int prev1,
                                                   double& best0, int& arg0, double& best1, int& arg1) {
    constexpr int hw = HW, W = 2 * HW + 1;
    constexpr int NCHW = (31 + HW) / 8 + (HW + 7) / 8 + 1;   // chunks that can touch the band of some destination of a warp
    static_assert(NCHW <= 32, "one lane per window chunk");
    const double NEG_INF = -INFINITY;
    const double LOGTINY = p.log_tiny;
    const double* Vc0 = &s.V[cur][0][VT_HALO];
    const double* Vc1 = &s.V[cur][1][VT_HALO];
    const unsigned char* rvc = &s.rowvar[VT_HALO];
    auto lt_at = [&](int var, int sel, int o) -> double {   // 0 <= o < W
        return (var < nsv) ? s.lt[var][sel][VT_LT_PAD + o] : __ldg(p.lt_variants + (static_cast<long long>(var) * 2 + sel) * W + o);
    };

    // ---- lower bounds: the candidates from this destination's own bin, and from the source that won in the previous frame
    // (decoded paths move slowly, so that real candidate is usually (near) optimal and prunes almost every chunk)
    double L0 = NEG_INF, L1;
    {
        const int var = rvc[d];
        const double ls = lt_at(var, 0, hw), lx = lt_at(var, 1, hw);
        const double x0 = Vc0[d], x1 = Vc1[d];
        L1 = fmax(x0 + lx, x1 + ls);
        if (BOTH) L0 = fmax(x0 + ls, x1 + lx);
        const int sv1 = prev1 >= n, bs1 = prev1 - sv1 * n, o1 = d - bs1 + hw;
        if (o1 >= 0 && o1 < W) L1 = fmax(L1, (sv1 ? Vc1 : Vc0)[bs1] + lt_at(rvc[bs1], 1 - sv1, o1));
        if (BOTH) {
            const int sv0 = prev0 >= n, bs0 = prev0 - sv0 * n, o0 = d - bs0 + hw;
            if (o0 >= 0 && o0 < W) L0 = fmax(L0, (sv0 ? Vc1 : Vc0)[bs0] + lt_at(rvc[bs0], sv0, o0));
        }
    }
    const double Lmin1 = lower_bound_from_hi(__reduce_max_sync(0xffffffffu, live ? static_cast<unsigned>(__double2hiint(L1)) : 0u));
    double Lmin0 = INFINITY;
    if (BOTH) Lmin0 = lower_bound_from_hi(__reduce_max_sync(0xffffffffu, live ? static_cast<unsigned>(__double2hiint(L0)) : 0u));

    best0 = NEG_INF; best1 = NEG_INF;
    arg0 = 0; arg1 = 0;
    const int c_lo = 4 * warp - (HW + 7) / 8;       // floor((32 warp - hw) / 8)
#pragma unroll 1
    for (int sv = 0; sv < 2; ++sv) {                // source block: 0 voiced, 1 unvoiced (ascending state index); one copy of the code
        const double* Vc = sv == 0 ? Vc0 : Vc1;
        const int sel1 = 1 - sv;                    // table for an unvoiced destination: switch from voiced, same from unvoiced
        const int sel0 = sv;                        // table for a voiced destination
        const int kbase = sv * n;
        // ---- out-of-band sources
        unsigned hmin = lane < n_warps ? s.seg_hi[cur][sv][lane] : 0xFFFFFFFFu;
        hmin = __reduce_min_sync(0xffffffffu, hmin);
        const double oob_ub = upper_bound_from_hi(hmin) + LOGTINY;
        const bool oob_possible = (oob_ub >= Lmin1) || (BOTH && oob_ub >= Lmin0);
        VA g{NEG_INF, 0x7fffffff};
        double oob = NEG_INF;
        if (oob_possible) {
            g = block_leftmost_max(Vc, n, lane);
            oob = g.v + LOGTINY;
            if (g.a < d - hw) {   // lower indices than the band
                if (oob > best1) { best1 = oob; arg1 = kbase + g.a; }
                if (BOTH && oob > best0) { best0 = oob; arg0 = kbase + g.a; }
            }
        }
        // ---- prefilter: lane r holds window chunk c_lo + r
        unsigned mask;
        {
            const int r = lane < NCHW ? lane : 0;
            const double m = s.M[cur][sv][c_lo + r + VT_CHUNK_PAD];
            bool keep = m + s.ubrmax[sel1][r] >= Lmin1;
            if (BOTH) keep = keep || (m + s.ubrmax[sel0][r] >= Lmin0);
            mask = __ballot_sync(0xffffffffu, keep && lane < NCHW);
        }
        while (mask) {
            const int r = __ffs(mask) - 1;
            mask &= mask - 1;
            const int c = c_lo + r;
            const double m = s.M[cur][sv][c + VT_CHUNK_PAD];
            const int ohi = d + hw - VT_CHUNK * c;   // offset of the chunk's first source in this destination's band
            const double bd1 = m + s.ubr[sel1][VT_UBR_PAD + ohi];
            bool need = live && (bd1 >= L1) && (bd1 > best1);
            if (BOTH) {
                const double bd0 = m + s.ubr[sel0][VT_UBR_PAD + ohi];
                need = need || (live && (bd0 >= L0) && (bd0 > best0));
            }
            if (!__any_sync(0xffffffffu, need)) continue;
            // ---- exact evaluation of the chunk's 8 sources for every lane (a lane that did not need it loses nothing:
            // the sources are real candidates, visited in ascending order)
            const int b0 = VT_CHUNK * c;
            const unsigned long long vars = *reinterpret_cast<const unsigned long long*>(&rvc[b0]);
            // one destination voicing and half a chunk at a time: 4 candidates live, not 16
            auto eval = [&](const int sel, double& best, int& arg) {
#pragma unroll 1
                for (int h = 0; h < VT_CHUNK; h += 4) {
                    const double2 xa = *reinterpret_cast<const double2*>(&Vc[b0 + h]);
                    const double2 xb = *reinterpret_cast<const double2*>(&Vc[b0 + h + 2]);
                    double cc[4] = {xa.x, xa.y, xb.x, xb.y};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int var = static_cast<int>((vars >> (8 * (h + j))) & 0xff);   // warp-uniform
                        const int o = ohi - h - j;
                        double tv;
                        if (var < nsv) {
                            tv = s.lt[var][sel][VT_LT_PAD + o];
                        } else {   // truncated edge row: global table, range-checked
                            const bool in = static_cast<unsigned>(o) < static_cast<unsigned>(W);
                            tv = in ? __ldg(p.lt_variants + (static_cast<long long>(var) * 2 + sel) * W + o) : NEG_INF;
                        }
                        cc[j] += tv;
                    }
                    // leftmost-max tournament over the 4 candidates, then against the running best (lower indices)
                    int i0 = h, i2 = h + 2;
                    take_later(cc[0], i0, cc[1], h + 1);
                    take_later(cc[2], i2, cc[3], h + 3);
                    take_later(cc[0], i0, cc[2], i2);
                    if (cc[0] > best) { best = cc[0]; arg = kbase + b0 + i0; }
                }
            };
            eval(sel1, best1, arg1);
            if (BOTH) eval(sel0, best0, arg0);
        }
        if (oob_possible && g.a > d + hw) {  // higher indices than the band
            if (oob > best1) { best1 = oob; arg1 = kbase + g.a; }
            if (BOTH && oob > best0) { best0 = oob; arg0 = kbase + g.a; }
        }
    }
}

// HW: half width of the transition band (compile time so the loops unroll).
template <int HW, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
viterbi_forward_kernel(const aegis_viterbi_params p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    VitSmem& s = *reinterpret_cast<VitSmem*>(smem_raw);
    const int clip = blockIdx.x;
    const int b = threadIdx.x, lane = b & 31, warp = b >> 5;
    constexpr int hw = HW, W = 2 * HW + 1;
    constexpr int NCHW = (31 + HW) / 8 + (HW + 7) / 8 + 1;
    static_assert(HW + 7 <= VT_HALO && (HW + 7) / 8 <= VT_CHUNK_PAD && HW <= VT_MAX_HW, "halo / padding too small");
    static_assert(31 + HW + ((8 - HW % 8) % 8) + HW <= W + VT_LT_PAD - 3 && VT_UBR_PAD >= 31, "table padding too small");
    const int n = p.n_pitch_bins, T = p.n_frames;
    const int n_warps = (n + 31) >> 5;
    const int nsv = min(p.n_interior_variants, VT_SMEM_VARIANTS);
    const double NEG_INF = -INFINITY;
    const double LOGTINY = p.log_tiny;
    const bool live = b < n;

    // ---- one-time shared set-up
    for (int i = b; i < 2 * 2 * (VT_MAX_BINS + 2 * VT_HALO); i += blockDim.x) (&s.V[0][0][0])[i] = NEG_INF;
    for (int i = b; i < 2 * 2 * VT_N_CHUNKS; i += blockDim.x) (&s.M[0][0][0])[i] = NEG_INF;
    for (int i = b; i < 2 * VT_MAX_BINS; i += blockDim.x) (&s.obs_lp[0][0])[i] = LOGTINY;
    for (int i = b; i < VT_SMEM_VARIANTS * 2 * VT_LT_PITCH; i += blockDim.x) (&s.lt[0][0][0])[i] = NEG_INF;
    for (int i = b; i < 2 * VT_UBR_SIZE; i += blockDim.x) (&s.ubr[0][0])[i] = NEG_INF;
    for (int i = b; i < VT_N_CHUNKS; i += blockDim.x) s.cdub[i] = 0.0;
    for (int i = b; i < 2 * 2 * VT_MAX_WARPS; i += blockDim.x) (&s.seg_hi[0][0][0])[i] = 0xFFF00000u;
    for (int i = b; i < VT_MAX_BINS + 2 * VT_HALO; i += blockDim.x) {
        const int src = i - VT_HALO;
        s.rowvar[i] = (src >= 0 && src < n) ? static_cast<unsigned char>(__ldg(p.row_variant + src)) : 0;
    }
    __syncthreads();
    for (int i = b; i < nsv * 2 * W; i += blockDim.x) {
        const int var = i / (2 * W), rem = i - var * 2 * W;
        s.lt[var][rem / W][VT_LT_PAD + rem % W] = __ldg(p.lt_variants + i);
    }
    __syncthreads();
    // Chunk upper bounds.  base[sel][o] = max over the interior variants held in shared memory (they differ in
    // the last ulp); ubr[sel][q] = max_{j<8} base[sel][q-j].  A source row with any other variant (the truncated
    // edge rows are up to log 2 larger) carries its own excess dub = max_{sel,o}(lt_row - base) + margin; a chunk's
    // bound adds the largest excess of its 8 rows (cdub): max_chunk(V) + cdub + ubr bounds every candidate of the chunk.
    for (int i = b; i < 2 * (W + VT_CHUNK - 1); i += blockDim.x) {
        const int sel = i / (W + VT_CHUNK - 1), q = i - sel * (W + VT_CHUNK - 1);
        double m = NEG_INF;
        for (int j = 0; j < VT_CHUNK; ++j) {
            const int o = q - j;
            if (o < 0 || o >= W) continue;
            for (int var = 0; var < nsv; ++var) m = fmax(m, s.lt[var][sel][VT_LT_PAD + o]);
        }
        s.ubr[sel][VT_UBR_PAD + q] = m;
    }
    __syncthreads();
    if (b < 2 * 32) {   // prefilter table: window chunk r against the 32 destinations of a warp (the same for every warp)
        const int sel = b >> 5, r = b & 31;
        double m = NEG_INF;
        if (r < NCHW) {
            constexpr int off0 = 8 * ((HW + 7) / 8);   // 32 warp - 8 c_lo
            for (int l = 0; l < 32; ++l) m = fmax(m, s.ubr[sel][VT_UBR_PAD + l + hw + off0 - VT_CHUNK * r]);
        }
        s.ubrmax[sel][r] = m;
    }
    if (live) {
        const int var = __ldg(p.row_variant + b);
        if (var >= nsv) {
            double ex = 0.0;
            for (int sel = 0; sel < 2; ++sel)
                for (int o = 0; o < W; ++o) {
                    const double v = __ldg(p.lt_variants + (static_cast<long long>(var) * 2 + sel) * W + o);
                    double base = NEG_INF;
                    for (int iv = 0; iv < nsv; ++iv) base = fmax(base, s.lt[iv][sel][VT_LT_PAD + o]);
                    if (v > NEG_INF) ex = fmax(ex, v - base);
                }
            const double dub = ex + 1e-6;  // margin >> any rounding of the bound arithmetic (|V| < 1e10)
            // cdub = max over the chunk's rows: values are non-negative, so the integer order of the bit patterns is theirs
            atomicMax(reinterpret_cast<unsigned long long*>(&s.cdub[(b >> 3) + VT_CHUNK_PAD]),
                      static_cast<unsigned long long>(__double_as_longlong(dub)));
        }
    }
    __syncthreads();

    const long long f0idx = static_cast<long long>(clip) * T;
    const unsigned short* __restrict__ cbin = p.cand_bin + f0idx * p.max_cand;
    const double* __restrict__ cprob = p.cand_prob + f0idx * p.max_cand;
    const int* __restrict__ ccnt = p.cand_count + f0idx;
    const double* __restrict__ vprob = p.voiced_prob + f0idx;
    unsigned short* __restrict__ bp_out = p.backptr + f0idx * (2 * n);

    if (b == 0 && T > 0) s.lp_u[0] = log((1.0 - __ldg(vprob)) / static_cast<double>(n) + DBL_MIN);
    {   // scatter frame 0 observations
        const int cnt = T > 0 ? min(__ldg(ccnt), p.max_cand) : 0;
        if (b < cnt) s.obs_lp[0][cbin[b]] = log(cprob[b] + DBL_MIN);
    }
    __syncthreads();

    double vnew0 = NEG_INF, vnew1 = NEG_INF;   // this bin's voiced / unvoiced value
    int prev0 = b, prev1 = n + b;               // winning sources of the previous frame (temporal coherence)
    for (int t = 0; t < T; ++t) {
        const int cur = t & 1, nxt = cur ^ 1;
        int ncnt = 0, nbin = 0;   // prefetch the next frame's sparse observation
        double nprob = 0.0;
        if (t + 1 < T) {
            ncnt = min(__ldg(ccnt + t + 1), p.max_cand);
            if (b < ncnt) {
                nbin = __ldg(cbin + static_cast<long long>(t + 1) * p.max_cand + b);
                nprob = __ldg(cprob + static_cast<long long>(t + 1) * p.max_cand + b);
            }
        }
        const double lp_u = s.lp_u[cur];
        double lp_v = LOGTINY;
        if (live) {
            lp_v = s.obs_lp[cur][b];
            s.obs_lp[cur][b] = LOGTINY;  // reset for frame t+2
        }
        const bool is_cand = live && lp_v > LOGTINY;
        bool have0;   // does this warp publish voiced values this frame (warp-uniform)

        if (t == 0) {
            // voiced states without a candidate start at 2 log(tiny): dominated by the unvoiced twin (>= log(tiny) + log(1/n))
            vnew0 = is_cand ? lp_v + LOGTINY : NEG_INF;   // log(p_init = 0 + tiny)
            vnew1 = lp_u + p.log_init_unvoiced;           // log(1/n + tiny)
            have0 = __any_sync(0xffffffffu, is_cand);
        } else {
            // dominance holds for the voiced states of this frame iff the unvoiced observation is well above log(tiny)
            const bool dense_v = !(lp_u > LOGTINY + 10.0);
            have0 = dense_v || __any_sync(0xffffffffu, is_cand);
            double best0, best1;
            int arg0, arg1;
            if (have0) viterbi_frame_step<HW, true>(s, p, cur, b, lane, warp, live, n, nsv, n_warps, prev0, prev1, best0, arg0, best1, arg1);
            else viterbi_frame_step<HW, false>(s, p, cur, b, lane, warp, live, n, nsv, n_warps, prev0, prev1, best0, arg0, best1, arg1);
            unsigned short* row = bp_out + static_cast<long long>(t) * (2 * n);
            vnew0 = NEG_INF;
            if (live) {
                vnew1 = lp_u + best1;
                row[n + b] = static_cast<unsigned short>(arg1);
                prev1 = arg1;
                if (have0) {
                    prev0 = arg0;
                    if (dense_v || is_cand) {
                        vnew0 = lp_v + best0;
                        row[b] = static_cast<unsigned short>(arg0);
                    }
                }
            }
        }

        // publish V[t], its chunk maxima and the per-warp high-word bound
        if (live) {
            s.V[nxt][0][VT_HALO + b] = vnew0;
            s.V[nxt][1][VT_HALO + b] = vnew1;
        }
#pragma unroll
        for (int v = 0; v < 2; ++v) {
            if (v == 0 && !have0) {   // nothing but -inf in this warp's voiced states
                if ((lane & 7) == 0) s.M[nxt][0][(b >> 3) + VT_CHUNK_PAD] = NEG_INF;
                if (lane == 0) s.seg_hi[nxt][0][warp] = 0xFFF00000u;
                continue;
            }
            const double xv = live ? (v == 0 ? vnew0 : vnew1) : NEG_INF;
            double cm = xv;
            cm = fmax(cm, __shfl_xor_sync(0xffffffffu, cm, 1));
            cm = fmax(cm, __shfl_xor_sync(0xffffffffu, cm, 2));
            cm = fmax(cm, __shfl_xor_sync(0xffffffffu, cm, 4));
            if ((lane & 7) == 0) s.M[nxt][v][(b >> 3) + VT_CHUNK_PAD] = cm + s.cdub[(b >> 3) + VT_CHUNK_PAD];
            const unsigned hmin = __reduce_min_sync(0xffffffffu, static_cast<unsigned>(__double2hiint(xv)));
            if (lane == 0) s.seg_hi[nxt][v][warp] = hmin;
        }
        // logs of the NEXT frame's observations, one copy of the (long) double-precision log: pass 0 the candidates
        // (threads below the candidate count), pass 1 the bin-independent unvoiced observation (the last thread)
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
            const bool mine = pass == 0 ? (b < ncnt) : (b == static_cast<int>(blockDim.x) - 1 && t + 1 < T);
            if (!__any_sync(0xffffffffu, mine)) continue;
            const double a = pass == 0 ? nprob + DBL_MIN : (1.0 - __ldg(vprob + min(t + 1, T - 1))) / static_cast<double>(n) + DBL_MIN;
            const double v = log(a);
            if (mine) {
                if (pass == 0) s.obs_lp[nxt][nbin] = v;
                else s.lp_u[nxt] = v;
            }
        }
        __syncthreads();
    }
    if (live && T > 0) {
        double* fv = p.final_value + static_cast<long long>(clip) * (2 * n);
        fv[b] = vnew0;
        fv[n + b] = vnew1;
    }
}

// back-trace: one thread per clip (latency bound; all clips walk in parallel).  States the forward pass skipped
// (dominated: final value -inf, no back-pointer) are never visited.
__global__ void __launch_bounds__(64)
viterbi_backtrace_kernel(const aegis_viterbi_params p) {
    const int clip = blockIdx.x * blockDim.x + threadIdx.x;
    if (clip >= p.n_clips) return;
    const int n = p.n_pitch_bins, T = p.n_frames;
    if (T <= 0) return;
    const double* fv = p.final_value + static_cast<long long>(clip) * (2 * n);
    int st = 0;
    double best = fv[0];
    for (int j = 1; j < 2 * n; ++j) {
        const double v = fv[j];
        if (v > best) { best = v; st = j; }
    }
    const long long base = static_cast<long long>(clip) * T;
    const unsigned short* bp = p.backptr + base * (2 * n);
    for (int t = T - 1; t >= 0; --t) {
        p.states[base + t] = static_cast<unsigned short>(st);
        const bool voiced = st < n;
        p.voiced_flag[base + t] = voiced ? 1 : 0;
        p.f0[base + t] = voiced ? __ldg(p.freqs + st) : p.fill_value;
        if (t > 0) st = bp[static_cast<long long>(t) * (2 * n) + st];
    }
}

}  // namespace aegis

extern "C" int aegis_viterbi(const aegis_viterbi_params* p, void* stream) {
    using namespace aegis;
    AEGIS_REQUIRE(p != nullptr, "aegis_viterbi: null params");
    AEGIS_REQUIRE(p->n_pitch_bins >= 2 && p->n_pitch_bins <= VT_MAX_BINS, "aegis_viterbi: n_pitch_bins=%d unsupported (<= %d)", p->n_pitch_bins, VT_MAX_BINS);
    AEGIS_REQUIRE(p->half_width >= 1 && p->half_width <= VT_MAX_HW, "aegis_viterbi: half_width=%d unsupported (<= %d)", p->half_width, VT_MAX_HW);
    AEGIS_REQUIRE(p->n_variants >= 1 && p->n_variants <= 255 && p->n_interior_variants >= 1, "aegis_viterbi: bad variant counts");
    AEGIS_REQUIRE(p->max_cand >= 1 && p->max_cand <= p->n_pitch_bins, "aegis_viterbi: max_cand must be 1..n_pitch_bins");
    AEGIS_REQUIRE(p->cand_bin && p->cand_prob && p->cand_count && p->voiced_prob && p->lt_variants && p->row_variant && p->freqs,
                  "aegis_viterbi: inputs missing");
    AEGIS_REQUIRE(p->backptr && p->final_value && p->states && p->f0 && p->voiced_flag, "aegis_viterbi: outputs missing");
    if (p->n_clips == 0 || p->n_frames == 0) return 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int block = ((p->n_pitch_bins + 31) / 32) * 32;
    const bool small = block <= 448;
    void (*kern)(const aegis_viterbi_params) = nullptr;
    switch (p->half_width) {   // pYIN's band: 50 bins at 22.05 kHz / hop 512, 25 at 44.1 kHz
        case 50: kern = small ? viterbi_forward_kernel<50, 448, 2> : viterbi_forward_kernel<50, 512, 2>; break;
        case 25: kern = small ? viterbi_forward_kernel<25, 448, 2> : viterbi_forward_kernel<25, 512, 2>; break;
        case 12: kern = small ? viterbi_forward_kernel<12, 448, 2> : viterbi_forward_kernel<12, 512, 2>; break;
        default: set_error("aegis_viterbi: half_width=%d has no compiled kernel (12, 25, 50)", p->half_width); return 1;
    }
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(VitSmem)));
    if (e != cudaSuccess) {
        set_error("aegis_viterbi: cannot reserve %zu B shared memory: %s", sizeof(VitSmem), cudaGetErrorString(e));
        return 2;
    }
    kern<<<p->n_clips, block, sizeof(VitSmem), st>>>(*p);
    if (int rc = check_launch("aegis_viterbi(forward)")) return rc;
    viterbi_backtrace_kernel<<<(p->n_clips + 63) / 64, 64, 0, st>>>(*p);
    return check_launch("aegis_viterbi(backtrace)");
}
