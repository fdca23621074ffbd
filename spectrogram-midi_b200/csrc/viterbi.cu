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
// bin, the 32 destinations of a "window" walk the same source chunks together), so there is no
// divergence.  The (window, destination voicing) pairs of a frame are TASKS that the warps of the CTA
// pull from a shared-memory counter, expensive ones first (voiced destinations, edge windows), so
// the per-frame barrier waits for a balanced load; CTAs are small (7 warps for 441 bins) and four of
// them share an SM, so one clip's barrier is hidden behind three other clips:
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
//   4. COLLAPSE: in a frame whose unvoiced observation is (about) log(tiny) -- voiced_prob == 1, most frames of a clean
//      note -- every state without a candidate carries an observation <= lp_u, so its value is at most
//      lp_u + max(V[t-1]) + LTMAX (LTMAX = the largest log transition) and its candidate for ANY destination of the next
//      frame at most that + LTMAX, while a candidate state c offers at least V[t][c] + log(tiny) to every destination.
//      When  lower_bound(V[t][c]) + log(tiny) > lp_u + upper_bound(max V[t-1]) + 2 LTMAX + 1e-6  for some c, no state
//      without a candidate can win or tie anywhere in frame t+1, nor be the final state: the frame computes the voiced
//      destinations of the windows that hold candidates and nothing else (all other values -inf, no back-pointers).
//      The bounds use only V[t-1] and the candidate lists of frames t-1 and t, so every warp takes the same decision.
//   5. COLLAPSE RUN: after a collapsed frame the decoder state is just that frame's (<= 32) candidate states.  While the
//      following frames collapse too, ONE warp decodes them alone on the candidate lists (lane = candidate): no V images,
//      no task queue, no CTA barrier; the other warps wait once.  The run ends at the first frame that does not collapse;
//      the warp then rebuilds what the general frame code expects (see `collapse_run` in the kernel).
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
#include <cstdlib>
#include "common.cuh"

namespace aegis {

#ifndef VT_RUN_UNROLL
#define VT_RUN_UNROLL 4      // sources of a collapse-run frame evaluated together (1, 2 or 4: 32 is a multiple)
#endif
#ifndef VT_RUN_PRECHECK
#define VT_RUN_PRECHECK 0   // A/B timed: the test costs every frame more than the failed attempts it saves
#endif

#ifndef VT_MAX_BINS_DEF
#define VT_MAX_BINS_DEF 512
#endif
constexpr int VT_MAX_BINS = VT_MAX_BINS_DEF;
constexpr int VT_HALO = 64;            // V halo: sources outside [0, n) read -inf; needs half_width + 7 <= 64
constexpr int VT_MAX_HW = 50;
constexpr int VT_MAX_W = 2 * VT_MAX_HW + 1;
constexpr int VT_SMEM_VARIANTS = 3;    // interior row variants kept in shared memory (librosa's tables have 3; more are served like edge rows)
constexpr int VT_MAX_WIN = VT_MAX_BINS / 32;   // windows of 32 destination bins
constexpr int VT_CHUNK = 8;            // sources are pruned in aligned chunks of 8 bins
constexpr int VT_CHUNK_PAD = 8;        // chunk indices -8 .. (512/8 + 8)
constexpr int VT_N_CHUNKS = VT_MAX_BINS / VT_CHUNK + 2 * VT_CHUNK_PAD;
constexpr int VT_LT_PAD = 40;          // table offsets reach -38 .. W + 37 for lanes outside a chunk's band
constexpr int VT_LT_PITCH = VT_MAX_W + 2 * VT_LT_PAD + 1;   // odd number of doubles
constexpr int VT_UBR_PAD = 32;         // chunk offsets reach -31 .. W + 37
constexpr int VT_UBR_SIZE = VT_MAX_W + VT_UBR_PAD + 40;
constexpr unsigned VT_EDGE_BIT = 0x8000u;   // rowoff: the row's table lives in global memory, low 15 bits = variant * 2 * W

struct VitSmem {
    double V[2][2][VT_MAX_BINS + 2 * VT_HALO];     // [ping][voicing][halo | bins | halo], halo = -inf
    double M[2][2][VT_N_CHUNKS];                   // max of V over each aligned chunk of 8 bins + cdub (-inf outside)
    double obs_lp[2][VT_MAX_BINS];                 // log(obs + tiny) of the voiced states
    double lt[VT_SMEM_VARIANTS][2][VT_LT_PITCH];   // interior transition variants [variant][same|switch][pad + offset], -inf padding
    double ubr[2][VT_UBR_SIZE];                    // [same|switch][pad + q]: max over the interior variants and over offsets q-7..q
    double ubrmax[2][32];                          // [same|switch][r]: max over the 32 destinations of a window of ubr for window chunk r
    double cdub[VT_N_CHUNKS];                      // per chunk: how far its (edge) rows' tables exceed the interior ones
    double lp_u[2];                                // log-observation of the unvoiced states of frame t (same for all bins)
    // per source row: byte offset of its table inside lt[][0][] (interior rows), or VT_EDGE_BIT | element offset in lt_variants
    alignas(16) unsigned short rowoff[VT_MAX_BINS + 2 * VT_HALO];
    unsigned seg_hi[2][2][VT_MAX_WIN];             // per window: smallest high word of V (= upper bound of the window's maximum)
    unsigned short prev[2][VT_MAX_BINS];           // [destination voicing][bin]: winning source of the previous frame (temporal coherence)
    unsigned cand_windows[3];                      // bit w: window w holds a candidate bin in frame t (slot t % 3)
    int next_task[2];                              // task counter of frame t (slot t & 1)
    int cand_n[3];                                 // candidates of frame t (slot t % 3); the first 32 are listed below
    unsigned short cand_bin[3][32];
    double cand_lp[3][32];                         // log(prob + tiny) of the listed candidates
    double lt_max;                                 // largest finite log transition of the whole table
    int run_t;                                     // first frame a collapse run (warp 0 alone) left unprocessed
};

// table entry for (source row, same|switch, offset 0 <= o < W), any row
template <int W>
__device__ __forceinline__ double lt_entry(const VitSmem& s, const aegis_viterbi_params& p, int brow, int sel, int o) {
    const unsigned off = s.rowoff[VT_HALO + brow];
    if (!(off & VT_EDGE_BIT)) return *reinterpret_cast<const double*>(reinterpret_cast<const char*>(&s.lt[0][sel][VT_LT_PAD + o]) + off);
    return __ldg(p.lt_variants + (off & 0x7FFFu) + sel * W + o);
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
// max of two doubles that are never NaN (fmax's NaN handling costs two extra instructions per call)
__device__ __forceinline__ double dmax(double a, double b) { return b > a ? b : a; }
__device__ __forceinline__ double warp_max_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = dmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// Read-only tables (written once before the frame loop) are read through 32-bit shared addresses with an immediate
// offset: one LDS per value, no address arithmetic beyond one add.  Not volatile: the contents never change.
template <int IMM>
__device__ __forceinline__ double lds_table(unsigned addr) {
    double v;
    asm("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(addr), "n"(IMM));
    return v;
}

// The same task when the previous frame COLLAPSED: every state of V[t-1] is -inf except its (listed) voiced candidate
// states, so the dense argmax over all 2n sources reduces to those few, visited in ascending state index with a strict
// comparison -- in band with the row's table entry, out of band with log(tiny) -- which is the definition itself.
template <int HW>
__device__ __forceinline__ void viterbi_task_sparse(const VitSmem& s, const aegis_viterbi_params& p, const int cur, const int win,
                                                    const int dv, const int lane, const int slot_prev, const int n_prev,
                                                    double& best, int& arg) {
    constexpr int hw = HW, W = 2 * HW + 1;
    const int d = 32 * win + lane;
    best = -INFINITY;
    arg = 0;
#pragma unroll 1
    for (int k = 0; k < n_prev; ++k) {
        const int c = s.cand_bin[slot_prev][k];          // warp-uniform, ascending bins
        const double x = s.V[cur][0][VT_HALO + c];
        const int o = d - c + hw;
        const double tv = (o >= 0 && o < W) ? lt_entry<W>(s, p, c, dv, o) : p.log_tiny;   // table (voiced source -> dv) = dv
        const double v = x + tv;
        if (v > best) { best = v; arg = c; }
    }
}

// One task: the 32 destinations of window `win`, destination voicing `dv` (0 voiced, 1 unvoiced), frame t >= 1.
// Returns max_k (V[t-1][k] + log_trans[k -> destination]) and its first argmax for this lane's destination.
// Exactness of the pruning (everything is float64, first-index argmax as in numpy):
//  * a chunk of 8 sources is skipped for a destination when  chunk_max + max(lt over the chunk's offsets, all interior
//    variants) [+ the excess of edge rows, cdub]  is < L, the value of a real candidate of that destination (its own bin
//    or the source that won in the previous frame), or <= the running best of lower-index sources: floating-point
//    addition is monotone, so no source of the chunk can reach the maximum or win a tie.  The prefilter uses the same
//    bound maximised over the window's 32 destinations against a lower bound of the smallest L of the window.
//  * chunks that survive are evaluated exactly, in ascending source order (voiced block first), ties to the lower index.
//  * out-of-band sources of a block: let g be the block's leftmost maximum.  If g lies inside the band its in-band
//    candidate (>= V[g] - 14) beats every out-of-band one (<= V[g] - 708); otherwise g is the best out-of-band source
//    of its side and the other side can at best tie with a higher index.
template <int HW>
__device__ __forceinline__ void viterbi_task(const VitSmem& s, const aegis_viterbi_params& p, const int cur, const int win, const int dv,
                                             const int lane, const int n, const int n_win, double& best, int& arg) {
    constexpr int hw = HW, W = 2 * HW + 1;
    constexpr int NCHW = (31 + HW) / 8 + (HW + 7) / 8 + 1;   // chunks that can touch the band of some destination of a window
    static_assert(NCHW <= 32, "one lane per window chunk");
    const double NEG_INF = -INFINITY;
    const double LOGTINY = p.log_tiny;
    const int d = 32 * win + lane;
    const bool live = d < n;
    const double* Vc0 = &s.V[cur][0][VT_HALO];
    const unsigned short* roff = &s.rowoff[VT_HALO];
    constexpr int VROW = VT_MAX_BINS + 2 * VT_HALO;          // Vc1 = Vc0 + VROW
    auto lt_row = [&](int brow, int sel, int o) -> double { return lt_entry<W>(s, p, brow, sel, o); };

    // ---- lower bound L: the candidates from this destination's own bin, and from the source that won in the previous frame
    // (decoded paths move slowly, so that real candidate is usually (near) optimal and prunes almost every chunk)
    double L;
    {
        const int dr = min(d, n - 1);
        const double x0 = Vc0[d], x1 = Vc0[VROW + d];
        L = dmax(x0 + lt_row(dr, dv, hw), x1 + lt_row(dr, 1 - dv, hw));   // table (source voicing sv -> dv) = sv ^ dv
        const int pv = s.prev[dv][d];
        const int sv = pv >= n, bs = pv - sv * n, o = d - bs + hw;
        if (o >= 0 && o < W) L = dmax(L, Vc0[sv * VROW + bs] + lt_row(bs, sv ^ dv, o));
    }
    const double Lmin = lower_bound_from_hi(__reduce_max_sync(0xffffffffu, live ? static_cast<unsigned>(__double2hiint(L)) : 0u));

    best = NEG_INF;
    arg = 0;
    const int c_lo = 4 * win - (HW + 7) / 8;       // floor((32 win - hw) / 8)
    const int ohi0 = d + hw - VT_CHUNK * c_lo;
#pragma unroll 1
    for (int sv = 0; sv < 2; ++sv) {                // source block: 0 voiced, 1 unvoiced (ascending state index); one copy of the code
        const double* Vc = Vc0 + sv * VROW;
        const int sel = sv ^ dv;
        const int kbase = sv * n;
        // ---- out-of-band sources: can max(V) + log(tiny) reach any lane's lower bound at all?
        const unsigned seg = lane < n_win ? s.seg_hi[cur][sv][lane] : 0xFFFFFFFFu;
        const unsigned hmin = __reduce_min_sync(0xffffffffu, seg);
        const bool oob_possible = upper_bound_from_hi(hmin) + LOGTINY >= Lmin;
        int ga = 0x7fffffff;
        double oob = NEG_INF;
        if (oob_possible) {
            // exact leftmost maximum of the block: it lies in a window whose high-word bound equals the block's
            // (one window, barring values that agree to 20 bits); windows in ascending order, strict > keeps the first
            unsigned wm = __ballot_sync(0xffffffffu, seg == hmin);
            double gv = NEG_INF;
            while (wm) {
                const int w = __ffs(wm) - 1;
                wm &= wm - 1;
                const int i = 32 * w + lane;
                const double v = i < n ? Vc[i] : NEG_INF;
                const double m = warp_max_d(v);
                if (m > gv) {
                    gv = m;
                    ga = 32 * w + __ffs(__ballot_sync(0xffffffffu, v == m)) - 1;
                }
            }
            oob = gv + LOGTINY;
            if (ga < d - hw && oob > best) { best = oob; arg = kbase + ga; }   // lower indices than the band
        }
        // ---- prefilter: lane r holds window chunk c_lo + r
        const double* Mrow = &s.M[cur][sv][c_lo + VT_CHUNK_PAD];
        unsigned mask;
        double pb;   // lane r: bound of window chunk r that holds for all 32 destinations
        {
            const int r = lane < NCHW ? lane : 0;
            pb = lane < NCHW ? Mrow[r] + s.ubrmax[sel][r] : NEG_INF;
            mask = __ballot_sync(0xffffffffu, lane < NCHW && pb >= Lmin);   // (-inf >= -inf holds: the lane test is not redundant)
        }
        const double* ubr_l = &s.ubr[sel][VT_UBR_PAD + ohi0];      // this lane's bound for window chunk r: ubr_l[-8 r]
        // shared address of this lane's table entry for (variant 0, window chunk 0, source 0); chunk r, source j of a row
        // with table offset `off` sits at  + off - 64 r - 8 j
        const unsigned lt_l = static_cast<unsigned>(__cvta_generic_to_shared(&s.lt[0][sel][VT_LT_PAD + ohi0]));
        while (mask) {
            const int r = __ffs(mask) - 1;
            mask &= mask - 1;
            const double bd = Mrow[r] + ubr_l[-VT_CHUNK * r];
            if (!__any_sync(0xffffffffu, live && (bd >= L) && (bd > best))) continue;
            // ---- exact evaluation of the chunk's 8 sources for every lane (a lane that did not need it loses nothing:
            // the sources are real candidates, visited in ascending order)
            const int b0 = VT_CHUNK * (c_lo + r);
            const uint4 ro = *reinterpret_cast<const uint4*>(&roff[b0]);   // 8 table offsets (16 bit each), warp-uniform
            if (((ro.x | ro.y | ro.z | ro.w) & 0x80008000u) == 0u) {
                // all 8 rows are interior rows: tables in shared memory, -inf padded, no range checks
                const unsigned base = lt_l - 8 * VT_CHUNK * r;
#pragma unroll
                for (int h = 0; h < VT_CHUNK; h += 4) {   // half a chunk at a time: 4 candidates live
                    const double2 xa = *reinterpret_cast<const double2*>(&Vc[b0 + h]);
                    const double2 xb = *reinterpret_cast<const double2*>(&Vc[b0 + h + 2]);
                    const unsigned ra = h == 0 ? ro.x : ro.z, rb = h == 0 ? ro.y : ro.w;
                    double c0 = xa.x + (h == 0 ? lds_table<0>(base + (ra & 0xFFFFu)) : lds_table<-32>(base + (ra & 0xFFFFu)));
                    double c1 = xa.y + (h == 0 ? lds_table<-8>(base + (ra >> 16)) : lds_table<-40>(base + (ra >> 16)));
                    double c2 = xb.x + (h == 0 ? lds_table<-16>(base + (rb & 0xFFFFu)) : lds_table<-48>(base + (rb & 0xFFFFu)));
                    double c3 = xb.y + (h == 0 ? lds_table<-24>(base + (rb >> 16)) : lds_table<-56>(base + (rb >> 16)));
                    // leftmost-max tournament over the 4 candidates, then against the running best (lower indices)
                    int i0 = h, i2 = h + 2;
                    take_later(c0, i0, c1, h + 1);
                    take_later(c2, i2, c3, h + 3);
                    take_later(c0, i0, c2, i2);
                    if (c0 > best) { best = c0; arg = kbase + b0 + i0; }
                }
                if (mask) {   // chunks whose window-wide bound is <= every lane's running best cannot win (strict >) any more
                    const double bmin = lower_bound_from_hi(__reduce_max_sync(0xffffffffu, live ? static_cast<unsigned>(__double2hiint(best)) : 0u));
                    mask &= __ballot_sync(0xffffffffu, pb > bmin);
                }
            } else {
                // a chunk with truncated edge rows: their tables are in global memory (L1 resident), range-checked
                const int ohi = ohi0 - VT_CHUNK * r;    // offset of the chunk's first source in this destination's band
                const double* gt = p.lt_variants + sel * W + ohi;
                const unsigned base = lt_l - 8 * VT_CHUNK * r;
                const unsigned rr[4] = {ro.x, ro.y, ro.z, ro.w};
#pragma unroll
                for (int j = 0; j < VT_CHUNK; ++j) {
                    const unsigned off = (j & 1) ? (rr[j >> 1] >> 16) : (rr[j >> 1] & 0xFFFFu);   // warp-uniform
                    double tv;
                    if (!(off & VT_EDGE_BIT)) {
                        tv = *reinterpret_cast<const double*>(reinterpret_cast<const char*>(&s.lt[0][sel][VT_LT_PAD + ohi - j]) + off);
                    } else {
                        const bool in = static_cast<unsigned>(ohi - j) < static_cast<unsigned>(W);
                        tv = in ? __ldg(gt + (off & 0x7FFFu) - j) : NEG_INF;
                    }
                    const double v = Vc[b0 + j] + tv;
                    if (v > best) { best = v; arg = kbase + b0 + j; }   // ascending sources, strict >
                }
                (void)base;
            }
        }
        if (oob_possible && ga > d + hw && oob > best) { best = oob; arg = kbase + ga; }   // higher indices than the band
    }
}

// HW: half width of the transition band (compile time so the loops unroll).  NW warps per CTA.
template <int HW, int NW, int MINB>
__global__ void __launch_bounds__(32 * NW, MINB)
viterbi_forward_kernel(const aegis_viterbi_params p) {
    __shared__ VitSmem s;   // static (< 48 KB): every access is an LDS / STS with an immediate offset
    const int clip = blockIdx.x;
    const int b = threadIdx.x, lane = b & 31, warp = b >> 5;
    constexpr int hw = HW, W = 2 * HW + 1;
    constexpr int NCHW = (31 + HW) / 8 + (HW + 7) / 8 + 1;
    constexpr int NT = 32 * NW;
    static_assert(HW + 7 <= VT_HALO && (HW + 7) / 8 <= VT_CHUNK_PAD && HW <= VT_MAX_HW, "halo / padding too small");
    static_assert(VT_UBR_PAD >= 31 && VT_LT_PAD >= 38, "table padding too small");
    const int n = p.n_pitch_bins, T = p.n_frames;
    const int n_win = (n + 31) >> 5;
    const int nsv = min(p.n_interior_variants, VT_SMEM_VARIANTS);
    const double NEG_INF = -INFINITY;
    const double LOGTINY = p.log_tiny;

    // ---- one-time shared set-up
    for (int i = b; i < 2 * 2 * (VT_MAX_BINS + 2 * VT_HALO); i += NT) (&s.V[0][0][0])[i] = NEG_INF;
    for (int i = b; i < 2 * 2 * VT_N_CHUNKS; i += NT) (&s.M[0][0][0])[i] = NEG_INF;
    for (int i = b; i < 2 * VT_MAX_BINS; i += NT) (&s.obs_lp[0][0])[i] = LOGTINY;
    for (int i = b; i < VT_SMEM_VARIANTS * 2 * VT_LT_PITCH; i += NT) (&s.lt[0][0][0])[i] = NEG_INF;
    for (int i = b; i < 2 * VT_UBR_SIZE; i += NT) (&s.ubr[0][0])[i] = NEG_INF;
    for (int i = b; i < VT_N_CHUNKS; i += NT) s.cdub[i] = 0.0;
    for (int i = b; i < 2 * 2 * VT_MAX_WIN; i += NT) (&s.seg_hi[0][0][0])[i] = 0xFFF00000u;
    for (int i = b; i < VT_MAX_BINS + 2 * VT_HALO; i += NT) {
        const int src = i - VT_HALO;
        unsigned off = 0;   // rows outside [0, n) carry V = -inf: any table will do
        if (src >= 0 && src < n) {
            const int var = __ldg(p.row_variant + src);
            off = var < nsv ? static_cast<unsigned>(var * 2 * VT_LT_PITCH * sizeof(double)) : (VT_EDGE_BIT | static_cast<unsigned>(var * 2 * W));
        }
        s.rowoff[i] = static_cast<unsigned short>(off);
    }
    for (int i = b; i < VT_MAX_BINS; i += NT) {
        s.prev[0][i] = static_cast<unsigned short>(min(i, n - 1));
        s.prev[1][i] = static_cast<unsigned short>(n + min(i, n - 1));
    }
    if (b < 3) { s.cand_windows[b] = 0u; s.cand_n[b] = 0; }
    if (b < 2) s.next_task[b] = 0;
    if (b == 0) s.lt_max = NEG_INF;
    __syncthreads();
    {   // largest finite log transition (all entries are negative: the bit pattern shrinks as the value grows)
        double m = NEG_INF;
        for (int i = b; i < p.n_variants * 2 * W; i += NT) {
            const double v = __ldg(p.lt_variants + i);
            if (v > m) m = v;
        }
        atomicMin(reinterpret_cast<unsigned long long*>(&s.lt_max), static_cast<unsigned long long>(__double_as_longlong(m)));
    }
    for (int i = b; i < nsv * 2 * W; i += NT) {
        const int var = i / (2 * W), rem = i - var * 2 * W;
        s.lt[var][rem / W][VT_LT_PAD + rem % W] = __ldg(p.lt_variants + i);
    }
    __syncthreads();
    // Chunk upper bounds.  base[sel][o] = max over the interior variants held in shared memory (they differ in
    // the last ulp); ubr[sel][q] = max_{j<8} base[sel][q-j].  A source row with any other variant (the truncated
    // edge rows are up to log 2 larger) carries its own excess dub = max_{sel,o}(lt_row - base) + margin; a chunk's
    // bound adds the largest excess of its 8 rows (cdub): max_chunk(V) + cdub + ubr bounds every candidate of the chunk.
    for (int i = b; i < 2 * (W + VT_CHUNK - 1); i += NT) {
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
    if (b < 2 * 32) {   // prefilter table: window chunk r against the 32 destinations of a window (the same for every window)
        const int sel = b >> 5, r = b & 31;
        double m = NEG_INF;
        if (r < NCHW) {
            constexpr int off0 = 8 * ((HW + 7) / 8);   // 32 win - 8 c_lo
            for (int l = 0; l < 32; ++l) m = fmax(m, s.ubr[sel][VT_UBR_PAD + l + hw + off0 - VT_CHUNK * r]);
        }
        s.ubrmax[sel][r] = m;
    }
    for (int row = b; row < n; row += NT) {
        const int var = __ldg(p.row_variant + row);
        if (var >= nsv) {
            double ex = 0.0;
            for (int sel = 0; sel < 2; ++sel)
#pragma unroll 1
                for (int o = 0; o < W; ++o) {
                    const double v = __ldg(p.lt_variants + (static_cast<long long>(var) * 2 + sel) * W + o);
                    double base = NEG_INF;
                    for (int iv = 0; iv < nsv; ++iv) base = fmax(base, s.lt[iv][sel][VT_LT_PAD + o]);
                    if (v > NEG_INF) ex = fmax(ex, v - base);
                }
            const double dub = ex + 1e-6;  // margin >> any rounding of the bound arithmetic (|V| < 1e10)
            // cdub = max over the chunk's rows: values are non-negative, so the integer order of the bit patterns is theirs
            atomicMax(reinterpret_cast<unsigned long long*>(&s.cdub[(row >> 3) + VT_CHUNK_PAD]),
                      static_cast<unsigned long long>(__double_as_longlong(dub)));
        }
    }

    const long long f0idx = static_cast<long long>(clip) * T;
    const unsigned short* __restrict__ cbin = p.cand_bin + f0idx * p.max_cand;
    const double* __restrict__ cprob = p.cand_prob + f0idx * p.max_cand;
    const int* __restrict__ ccnt = p.cand_count + f0idx;
    const double* __restrict__ vprob = p.voiced_prob + f0idx;
    unsigned short* __restrict__ bp_out = p.backptr + f0idx * (2 * n);

    if (b == 0) s.lp_u[0] = log((1.0 - __ldg(vprob)) / static_cast<double>(n) + DBL_MIN);
    {   // scatter frame 0 observations
        const int cnt = min(__ldg(ccnt), p.max_cand);
        for (int i = b; i < cnt; i += NT) {
            const int bin = cbin[i];
            const double v = log(cprob[i] + DBL_MIN);
            s.obs_lp[0][bin] = v;
            atomicOr(&s.cand_windows[0], 1u << (bin >> 5));
            if (i < 32) { s.cand_bin[0][i] = static_cast<unsigned short>(bin); s.cand_lp[0][i] = v; }
        }
        if (b == 0) s.cand_n[0] = cnt;
    }
    __syncthreads();

    // publish one window of one voicing of V[t]: the values, their chunk maxima and the window's high-word bound
    auto publish = [&](const int nxt, const int v, const int win, const double xv) {
        const int d = 32 * win + lane;
        if (d < n) s.V[nxt][v][VT_HALO + d] = xv;
        double cm = xv;
        cm = dmax(cm, __shfl_xor_sync(0xffffffffu, cm, 1));
        cm = dmax(cm, __shfl_xor_sync(0xffffffffu, cm, 2));
        cm = dmax(cm, __shfl_xor_sync(0xffffffffu, cm, 4));
        const int ch = (d >> 3) + VT_CHUNK_PAD;
        if ((lane & 7) == 0) s.M[nxt][v][ch] = cm + s.cdub[ch];
        const unsigned hmin = __reduce_min_sync(0xffffffffu, static_cast<unsigned>(__double2hiint(xv)));
        if (lane == 0) s.seg_hi[nxt][v][win] = hmin;
    };
    auto publish_empty = [&](const int nxt, const int v, const int win) {   // a window whose states of voicing v are all dominated
        const int d = 32 * win + lane;
        if (d < n) s.V[nxt][v][VT_HALO + d] = NEG_INF;
        if ((lane & 7) == 0) s.M[nxt][v][(d >> 3) + VT_CHUNK_PAD] = NEG_INF;
        if (lane == 0) s.seg_hi[nxt][v][win] = 0xFFF00000u;
    };

    // COLLAPSE RUN.  After a collapsed frame the whole decoder state is the (<= 32) candidate states of that frame; while
    // the following frames collapse too, a frame is: for every candidate of frame t the strict first-index maximum over the
    // candidates of frame t-1 (the arithmetic of viterbi_task_sparse, destination = the candidate's own bin), the collapse
    // test on the same numbers, one back-pointer per candidate.  One WARP does that alone, lane = candidate, without the
    // shared-memory images of V, without task counter and without CTA barrier, while the other warps wait; it stops at the
    // first frame that does not collapse (or is the last one, or has no / too many candidates), rebuilds the images the
    // general frame code expects (V[t-1] with chunk maxima and window bounds, the observation of frame t, the candidate
    // lists of t-1 and t) and returns that frame's index.  Every collapse decision satisfies the bound of the file header
    // (with max V[t-1] itself in place of its high-word bound), so the decoded path is unchanged.
    auto collapse_run = [&](const int t0) -> int {
        constexpr unsigned FULL = 0xffffffffu;
        const int mc = p.max_cand;
        int t = t0;
        int Kp = s.cand_n[(t + 2) % 3];
        // previous frame, lane = candidate: value, and bin | table offset of its row << 16 (one shuffle brings both)
        int pq = s.rowoff[VT_HALO] << 16;
        double pv = NEG_INF;
        if (lane < Kp) {
            const int pb = s.cand_bin[(t + 2) % 3][lane];
            pq = pb | (s.rowoff[VT_HALO + pb] << 16);
            pv = s.V[t & 1][0][VT_HALO + pb];
        }
        int K = s.cand_n[t % 3];
        int cb = 0;
        double clp = NEG_INF;
        if (lane < min(K, 32)) {
            cb = s.cand_bin[t % 3][lane];
            clp = s.cand_lp[t % 3][lane];
        }
        double lpu = s.lp_u[t & 1];
        bool dense = !(lpu > LOGTINY + 10.0);
        // observations of the next two frames travel in registers (one DRAM round trip is longer than a frame of the run)
        int ak = 0, abin = 0, bk = 0, bbin = 0;
        double aprob = 0.0, avp = 0.0, bprob = 0.0, bvp = 0.0;
        auto fetch = [&](const int f, int& k, int& bin, double& prob, double& vp) {
            k = 0; bin = 0; prob = 0.0; vp = 0.0;
            if (f < T) {
                k = __ldg(ccnt + f);
                vp = __ldg(vprob + f);
                if (lane < mc) {
                    bin = __ldg(cbin + static_cast<long long>(f) * mc + lane);
                    prob = __ldg(cprob + static_cast<long long>(f) * mc + lane);
                }
            }
        };
        fetch(t + 1, ak, abin, aprob, avp);
        fetch(t + 2, bk, bbin, bprob, bvp);
        const double lt2 = 2.0 * s.lt_max;
        // one source: candidate k of frame t-1 against this lane's destination (lanes >= Kp hold -inf: they never win)
        auto source = [&](const int k, double& v, int& c, bool& in) {
            const double x = __shfl_sync(FULL, pv, k);
            const int q = __shfl_sync(FULL, pq, k);
            c = q & 0xFFFF;
            const unsigned off = static_cast<unsigned>(q) >> 16;
            const int o = cb - c + hw;
            in = static_cast<unsigned>(o) < static_cast<unsigned>(W);
            double tv = LOGTINY;
            if (in) {
                if (!(off & VT_EDGE_BIT)) tv = *reinterpret_cast<const double*>(reinterpret_cast<const char*>(&s.lt[0][0][VT_LT_PAD + o]) + off);
                else tv = __ldg(p.lt_variants + (off & 0x7FFFu) + o);
            }
            v = x + tv;
        };
        while (dense && K >= 1 && K <= 32 && t < T - 1) {
            const int cro = s.rowoff[VT_HALO + cb];   // needed when this frame has become the previous one
            double best = NEG_INF, inb = NEG_INF;
            int arg = 0;
#pragma unroll 1
            for (int k = 0; k < Kp; k += VT_RUN_UNROLL) {   // ascending bins, strict >: the first index wins a tie; VT_RUN_UNROLL sources in flight
                double v[VT_RUN_UNROLL];
                int c[VT_RUN_UNROLL];
                bool in[VT_RUN_UNROLL];
#pragma unroll
                for (int j = 0; j < VT_RUN_UNROLL; ++j) source(k + j, v[j], c[j], in[j]);
#pragma unroll
                for (int j = 0; j < VT_RUN_UNROLL; ++j) {
                    if (v[j] > best) { best = v[j]; arg = c[j]; }
                    if (in[j] && v[j] > inb) inb = v[j];
                }
            }
            const double lbmax = warp_max_d(lane < K ? clp + inb : NEG_INF);   // a real in-band candidate of a candidate state
            const double vmax = warp_max_d(pv);
            if (!(lbmax + LOGTINY > lpu + vmax + lt2 + 1e-6)) break;
            const bool has = lane < K && clp > LOGTINY;
            if (has) bp_out[static_cast<long long>(t) * (2 * n) + cb] = static_cast<unsigned short>(arg);
            if (lane < K) {
                s.prev[0][cb] = static_cast<unsigned short>(arg);
                if (t == t0) s.obs_lp[t & 1][cb] = LOGTINY;   // the only frame of the run whose observation was scattered
            }
            Kp = K; pq = cb | (cro << 16); pv = has ? clp + best : NEG_INF;
            ++t;
            K = min(ak, mc);
            cb = lane < K ? abin : 0;
            {
                const double l = log(aprob + DBL_MIN);
                clp = lane < K ? l : NEG_INF;
            }
            dense = avp == 1.0;   // <=> log((1 - vp) / n + tiny) <= log(tiny) + 10 for vp in [0, 1]
            lpu = LOGTINY;
            ak = bk; abin = bbin; aprob = bprob; avp = bvp;
            fetch(t + 2, bk, bbin, bprob, bvp);
        }
        const int pb = pq & 0xFFFF;
        if (t == t0) return t0;
        // ---- hand frame t to the general code: V[t-1] = the candidates of frame t-1, everything else -inf
        const int cur = t & 1;
        for (int w = 0; w < n_win; ++w) {
            publish_empty(cur, 0, w);
            publish_empty(cur, 1, w);
        }
        __syncwarp();
        if (lane < Kp) {
            s.V[cur][0][VT_HALO + pb] = pv;
            s.cand_bin[(t + 2) % 3][lane] = static_cast<unsigned short>(pb);
        }
        unsigned pm = __reduce_or_sync(FULL, lane < Kp ? 1u << (pb >> 5) : 0u);
        __syncwarp();
        for (; pm; pm &= pm - 1) {
            const int w = __ffs(pm) - 1, d = 32 * w + lane;
            publish(cur, 0, w, d < n ? s.V[cur][0][VT_HALO + d] : NEG_INF);
        }
        const int Kt = min(__ldg(ccnt + t), mc);
        unsigned wm = 0u;
        for (int i = lane; i < Kt; i += 32) {
            const int bin = __ldg(cbin + static_cast<long long>(t) * mc + i);
            const double v = log(__ldg(cprob + static_cast<long long>(t) * mc + i) + DBL_MIN);
            s.obs_lp[cur][bin] = v;
            wm |= 1u << (bin >> 5);
            if (i < 32) { s.cand_bin[t % 3][i] = static_cast<unsigned short>(bin); s.cand_lp[t % 3][i] = v; }
        }
        wm = __reduce_or_sync(FULL, wm);
        if (lane == 0) {
            s.cand_n[(t + 2) % 3] = Kp;
            s.cand_n[t % 3] = Kt;
            s.cand_windows[t % 3] = wm;
            s.cand_windows[(t + 1) % 3] = 0u;
            s.lp_u[cur] = log((1.0 - __ldg(vprob + t)) / static_cast<double>(n) + DBL_MIN);
            s.next_task[0] = 0;
            s.next_task[1] = 0;
        }
        return t;
    };

    bool prev_collapsed = false;   // frame t-1 collapsed: V[t-1] holds nothing but its candidate states (warp-uniform, same in every warp)
    bool just_ran = false;         // the previous iteration was a collapse run: frame t is the one it could not take
    for (int t = 0; t < T; ++t) {
#ifndef VT_NO_RUN   // A/B builds only
        // (a frame with voiced_prob < 1 or without / with too many candidates cannot collapse: no attempt, no barrier)
#if VT_RUN_PRECHECK
        if (prev_collapsed && !just_ran && t < T - 1 && !(s.lp_u[t & 1] > LOGTINY + 10.0) && s.cand_n[t % 3] >= 1 && s.cand_n[t % 3] <= 32) {
#else
        if (prev_collapsed && !just_ran && t < T - 1) {
#endif
            if (warp == 0) {
                const int t1 = collapse_run(t);
                if (lane == 0) s.run_t = t1;
            }
            __syncthreads();
            const int t1 = s.run_t;
            if (t1 != t) {
                t = t1 - 1;
                just_ran = true;
                continue;
            }
        }
        just_ran = false;
#endif
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
        if (b == 0) {
            s.cand_windows[(t + 2) % 3] = 0u;   // last read in frame t-1, next filled at the end of frame t+1
            s.next_task[nxt] = 0;               // last used in frame t-1, next used in frame t+1
        }
        const double lp_u = s.lp_u[cur];
        unsigned short* bp_row = bp_out + static_cast<long long>(t) * (2 * n);
        double* fv = (t == T - 1) ? p.final_value + static_cast<long long>(clip) * (2 * n) : nullptr;
        if (t == 0) {
            // voiced states without a candidate start at 2 log(tiny): dominated by the unvoiced twin (>= log(tiny) + log(1/n))
            for (int win = warp; win < n_win; win += NW) {
                const int d = 32 * win + lane;
                double lp_v = LOGTINY;
                if (d < n) {
                    lp_v = s.obs_lp[0][d];
                    s.obs_lp[0][d] = LOGTINY;   // reset for frame 2
                }
                const double v0 = lp_v > LOGTINY ? lp_v + LOGTINY : NEG_INF;              // log(p_init = 0 + tiny)
                const double v1 = d < n ? lp_u + p.log_init_unvoiced : NEG_INF;           // log(1/n + tiny)
                publish(nxt, 0, win, v0);
                publish(nxt, 1, win, v1);
                if (fv && d < n) { fv[d] = v0; fv[n + d] = v1; }
            }
        } else {
            // Dominance holds for the voiced states of this frame iff the unvoiced observation is well above log(tiny):
            // then only the windows that hold a candidate compute voiced destinations.
            const bool dense_v = !(lp_u > LOGTINY + 10.0);
            const unsigned all_win = n_win >= 32 ? 0xFFFFFFFFu : ((1u << n_win) - 1u);
            const unsigned cmask = s.cand_windows[t % 3];
            // COLLAPSE test (see the file header): inputs are V[t-1] and the candidate lists of frames t-1 and t only
            bool collapse = false;
            if (dense_v) {
                const int K = s.cand_n[t % 3], Kp = s.cand_n[(t + 2) % 3];
                if (K >= 1 && K <= 32 && Kp <= 32) {
                    double lb = NEG_INF;
                    if (lane < K) {
                        const int c = s.cand_bin[t % 3][lane];
                        // a real candidate of (voiced, c): its own bin's unvoiced state, or a candidate state of frame t-1 in band
                        double L = s.V[cur][1][VT_HALO + c] + lt_entry<W>(s, p, c, 1, hw);
                        for (int k = 0; k < Kp; ++k) {
                            const int cp = s.cand_bin[(t + 2) % 3][k];
                            const int o = c - cp + hw;
                            if (o >= 0 && o < W) L = dmax(L, s.V[cur][0][VT_HALO + cp] + lt_entry<W>(s, p, cp, 0, o));
                        }
                        lb = s.cand_lp[t % 3][lane] + L;
                    }
                    const double lbmax = warp_max_d(lb);
                    const unsigned h = lane < n_win ? min(s.seg_hi[cur][0][lane], s.seg_hi[cur][1][lane]) : 0xFFFFFFFFu;
                    const double vmax_ub = upper_bound_from_hi(__reduce_min_sync(0xffffffffu, h));
                    collapse = lbmax + LOGTINY > lp_u + vmax_ub + 2.0 * s.lt_max + 1e-6;
#ifdef VT_NO_COLLAPSE   // A/B builds only
                    collapse = false;
#endif
                }
            }
            const bool sparse_sources = prev_collapsed && s.cand_n[(t + 2) % 3] <= 32;   // all of frame t-1's finite states are listed
            const bool all_voiced = dense_v && !collapse;       // every voiced destination is computed and kept
            const unsigned vmask = all_voiced ? all_win : cmask;
            const int n_voiced = __popc(vmask);
            const int n_tasks = n_voiced + (collapse ? 0 : n_win);
            // tasks 0 .. n_voiced-1: voiced destinations of the windows in vmask; then the unvoiced destinations of every
            // window, edge windows first (their tables sit in global memory: the longest tasks start first)
            while (true) {
                int task = 0;
                if (lane == 0) task = atomicAdd(&s.next_task[cur], 1);
                task = __shfl_sync(0xffffffffu, task, 0);
                if (task >= n_tasks) break;
                int win, dv;
                if (task < n_voiced) {
                    dv = 0;
                    unsigned m = vmask;   // the task-th set bit
                    for (int k = 0; k < task; ++k) m &= m - 1;
                    win = all_voiced ? task : __ffs(m) - 1;
                } else {
                    dv = 1;
                    const int k = task - n_voiced;
                    win = (k & 1) ? n_win - 1 - (k >> 1) : (k >> 1);
                }
                const int d = 32 * win + lane;
                const bool live = d < n;
                double best;
                int arg;
                if (sparse_sources) viterbi_task_sparse<HW>(s, p, cur, win, dv, lane, (t + 2) % 3, s.cand_n[(t + 2) % 3], best, arg);
                else viterbi_task<HW>(s, p, cur, win, dv, lane, n, n_win, best, arg);
                double vnew = NEG_INF;
                if (dv == 1) {
                    if (live) {
                        vnew = lp_u + best;
                        bp_row[n + d] = static_cast<unsigned short>(arg);
                        s.prev[1][d] = static_cast<unsigned short>(arg);
                    }
                    if (!((vmask >> win) & 1u)) {   // nobody computes this window's voiced states
                        publish_empty(nxt, 0, win);
                        if (fv && live) fv[d] = NEG_INF;
                    }
                } else {
                    if (live) {
                        const double lp_v = s.obs_lp[cur][d];
                        s.obs_lp[cur][d] = LOGTINY;   // reset for frame t+2
                        s.prev[0][d] = static_cast<unsigned short>(arg);
                        if (all_voiced || lp_v > LOGTINY) {
                            vnew = lp_v + best;
                            bp_row[d] = static_cast<unsigned short>(arg);
                        }
                    }
                }
                publish(nxt, dv, win, vnew);
                if (fv && live) fv[dv * n + d] = vnew;
            }
            prev_collapsed = collapse;
            if (collapse) {   // everything that was not computed is dominated: -inf, no back-pointer
                for (int win = warp; win < n_win; win += NW) {
                    const int d = 32 * win + lane;
                    publish_empty(nxt, 1, win);
                    if (fv && d < n) fv[n + d] = NEG_INF;
                    if (!((vmask >> win) & 1u)) {
                        publish_empty(nxt, 0, win);
                        if (fv && d < n) fv[d] = NEG_INF;
                    }
                }
            }
        }
        // logs of the NEXT frame's observations, one copy of the (long) double-precision log: pass 0 the candidates
        // (threads below the candidate count), pass 1 the bin-independent unvoiced observation (the last thread)
        if (t + 1 < T) {
#pragma unroll 1
            for (int pass = 0; pass < 2; ++pass) {
                const bool mine = pass == 0 ? (b < ncnt) : (b == NT - 1);
                if (!__any_sync(0xffffffffu, mine)) continue;
                const double a = pass == 0 ? nprob + DBL_MIN : (1.0 - __ldg(vprob + t + 1)) / static_cast<double>(n) + DBL_MIN;
                const double v = log(a);
                if (mine) {
                    if (pass == 0) {
                        s.obs_lp[nxt][nbin] = v;
                        atomicOr(&s.cand_windows[(t + 1) % 3], 1u << (nbin >> 5));
                        if (b < 32) { s.cand_bin[(t + 1) % 3][b] = static_cast<unsigned short>(nbin); s.cand_lp[(t + 1) % 3][b] = v; }
                    } else {
                        s.lp_u[nxt] = v;
                        s.cand_n[(t + 1) % 3] = ncnt;
                    }
                }
            }
            for (int i = b + NT; i < ncnt; i += NT) {   // more candidates than threads (wide lag ranges): rare
                const int bin = __ldg(cbin + static_cast<long long>(t + 1) * p.max_cand + i);
                s.obs_lp[nxt][bin] = log(__ldg(cprob + static_cast<long long>(t + 1) * p.max_cand + i) + DBL_MIN);
                atomicOr(&s.cand_windows[(t + 1) % 3], 1u << (bin >> 5));
            }
        }
        __syncthreads();
    }
}

// back-trace: one warp per clip.  The walk is a chain of dependent loads from a 2.3 GB array (1024 clips x 1292 frames x
// 882 states x 2 B): one DRAM round trip per frame when walked naively (1.24 ms).  A predecessor lies within half_width
// bins of its successor, in either voicing half, so the entries the next BT_DEPTH rows can possibly need form a cone
// around the current bin: the lanes fetch that cone in one burst of independent loads (one round trip), the walk then
// runs on shared memory.  A predecessor outside the cone (the log(0 + tiny) transitions are finite, so the forward pass
// may -- very rarely -- pick one) is read from global memory instead.  States the forward pass skipped (dominated: final
// value -inf, no back-pointer) are never visited.
constexpr int BT_DEPTH = 4;
constexpr int BT_WARPS = 4;
constexpr int BT_SPAN = 2 * (BT_DEPTH - 1) * VT_MAX_HW + 1;
__host__ __device__ constexpr int bt_loads_of_row(int k) { return (2 * k * VT_MAX_HW + 1 + 31) / 32; }   // per lane and half
__host__ __device__ constexpr int bt_loads_total() {
    int s = 0;
    for (int k = 0; k < BT_DEPTH; ++k) s += 2 * bt_loads_of_row(k);
    return s;
}
constexpr int BT_LOADS = bt_loads_total();

__global__ void __launch_bounds__(32 * BT_WARPS)
viterbi_backtrace_kernel(const aegis_viterbi_params p) {
    __shared__ unsigned short cone[BT_WARPS][BT_DEPTH][2][BT_SPAN];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int clip = blockIdx.x * BT_WARPS + warp;
    if (clip >= p.n_clips) return;
    const int n = p.n_pitch_bins, T = p.n_frames, hw = p.half_width;
    if (T <= 0) return;
    const double* fv = p.final_value + static_cast<long long>(clip) * (2 * n);
    // first maximum of the final values
    int st = 2 * n;
    double best = -INFINITY;
    for (int j = lane; j < 2 * n; j += 32) {
        const double v = fv[j];
        if (v > best || st == 2 * n) { best = v; st = j; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int os = __shfl_xor_sync(0xffffffffu, st, o);
        if (ov > best || (ov == best && os < st)) { best = ov; st = os; }
    }
    const long long base = static_cast<long long>(clip) * T;
    const unsigned short* bp = p.backptr + base * (2 * n);
    for (int t = T - 1; t >= 0; t -= BT_DEPTH) {
        const int b0 = st < n ? st : st - n;
        // row t - k: bins b0 - k hw .. b0 + k hw of both halves (row 0 has no back-pointers).  Every lane first issues
        // all of its loads, then parks them in shared memory: one round trip, not one per row
        unsigned short got[BT_LOADS];
        {
            int q0 = 0;
#pragma unroll
            for (int k = 0; k < BT_DEPTH; ++k) {
                const unsigned short* row = bp + static_cast<long long>(t - k) * (2 * n) + (b0 - k * hw);
                const int width = 2 * k * hw + 1;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
#pragma unroll
                    for (int q = 0; q < bt_loads_of_row(k); ++q) {
                        const int o = lane + 32 * q, bin = b0 - k * hw + o;
                        const bool live = o < width && t - k >= 1 && bin >= 0 && bin < n;
                        got[q0 + q] = live ? row[h * n + o] : static_cast<unsigned short>(0);
                    }
                    q0 += bt_loads_of_row(k);
                }
            }
            q0 = 0;
#pragma unroll
            for (int k = 0; k < BT_DEPTH; ++k) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
#pragma unroll
                    for (int q = 0; q < bt_loads_of_row(k); ++q)
                        if (lane + 32 * q < BT_SPAN) cone[warp][k][h][lane + 32 * q] = got[q0 + q];
                    q0 += bt_loads_of_row(k);
                }
            }
        }
        __syncwarp();
        // every lane walks the same chain; lane k keeps the state of frame t - k and writes it
        int mine = st;
#pragma unroll
        for (int k = 0; k < BT_DEPTH; ++k) {
            if (t - k < 0) break;
            if (lane == k) mine = st;
            if (t - k > 0) {
                const int h = st >= n, o = (st - h * n) - (b0 - k * hw);
                st = (o >= 0 && o <= 2 * k * hw) ? cone[warp][k][h][o] : bp[static_cast<long long>(t - k) * (2 * n) + st];
            }
        }
        __syncwarp();
        if (lane < BT_DEPTH && t - lane >= 0) {
            const bool voiced = mine < n;
            p.states[base + t - lane] = static_cast<unsigned short>(mine);
            p.voiced_flag[base + t - lane] = voiced ? 1 : 0;
            p.f0[base + t - lane] = voiced ? __ldg(p.freqs + mine) : p.fill_value;
        }
    }
}

template <int HW>
static int launch_forward(const aegis_viterbi_params* p, cudaStream_t st) {
    const int n_win = (p->n_pitch_bins + 31) / 32;
    void (*kern)(const aegis_viterbi_params) = nullptr;
    int nw = 0;
    // warps per CTA (tasks are pulled dynamically, so any count works): A/B timed through AEGIS_VT_WARPS
    static const int forced = []() { const char* e = getenv("AEGIS_VT_WARPS"); return e ? atoi(e) : 0; }();
    nw = forced ? forced : (n_win <= 14 ? 7 : 8);
#ifdef VT_FIVE   // A/B builds only (needs VT_MAX_BINS_DEF <= 448 for five CTAs per SM)
    if (nw == 5) kern = viterbi_forward_kernel<HW, 5, 5>;
    else
#endif
    if (nw == 6) kern = viterbi_forward_kernel<HW, 6, 4>;
    else if (nw == 7) kern = viterbi_forward_kernel<HW, 7, 4>;     // 441 bins (E2..C6): 7 warps, four clips per SM
    else { kern = viterbi_forward_kernel<HW, 8, 4>; nw = 8; }      // up to 512 bins
    static_assert(sizeof(VitSmem) <= 48 * 1024, "VitSmem must fit static shared memory");
    kern<<<p->n_clips, 32 * nw, 0, st>>>(*p);
    return check_launch("aegis_viterbi(forward)");
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
    int rc;
    switch (p->half_width) {   // pYIN's band: 50 bins at 22.05 kHz / hop 512, 25 at 44.1 kHz
        case 50: rc = launch_forward<50>(p, st); break;
        case 25: rc = launch_forward<25>(p, st); break;
        case 12: rc = launch_forward<12>(p, st); break;
        default: set_error("aegis_viterbi: half_width=%d has no compiled kernel (12, 25, 50)", p->half_width); return 1;
    }
    if (rc) return rc;
    viterbi_backtrace_kernel<<<(p->n_clips + BT_WARPS - 1) / BT_WARPS, 32 * BT_WARPS, 0, st>>>(*p);
    return check_launch("aegis_viterbi(backtrace)");
}
