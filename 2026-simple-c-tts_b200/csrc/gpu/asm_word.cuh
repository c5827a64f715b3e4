// asm_word.cuh -- the WORD_END op: silence trimming and phrase intonation.
#pragma once
#include "asm_common.cuh"

namespace ctts {

// ---------------------------------------------------------------- word end

// remove_silence_regions, ctts.c:1634, as a bitmask + scan + in-place compaction.
// Returns the new length.  `reg` = w + word_start, len = count - word_start.
__device__ uint32_t trim_region(const Smem& sm, const AsmArgs& A, uint32_t big, int16_t* reg, uint32_t len) {
    const int tid = threadIdx.x, lane = lane_id(), warp = warp_id();
    const uint32_t min_sil = A.prm.min_silence_samples;
    // 16-byte grid of the region: vector j holds region samples 8j - phase .. 8j - phase + 7
    const uint32_t phase = (uint32_t)((reinterpret_cast<uintptr_t>(reg) >> 1) & 7u);
    const int4* grid = reinterpret_cast<const int4*>(reg - phase);
    const uint32_t gvec = (phase + len + 7) >> 3;
    // max |x| over the region, |.| with the reference's int16 wrap (abs16(-32768) < 0 never wins).
    // The maximum of every 8-sample grid vector is kept (vm): a silent run long enough to be cut
    // must contain whole silent vectors, so most regions are cleared below without a second pass.
    const uint32_t wn = (len + 31) >> 5;
    const bool have_vm = 2 * wn + (gvec + 1) / 2 + 2 <= SCR_WORDS;
    uint16_t* vm = reinterpret_cast<uint16_t*>(sm.scratch + 2 * wn);   // after the bit mask arrays
    uint32_t pk2 = 0;
    for (uint32_t j = tid; j < gvec; j += ASM_THREADS) {
        int4 q = grid[j];
        const int i0 = 8 * (int)j - (int)phase;
        if (i0 < 0 || i0 + 8 > (int)len) {
            int16_t* e = reinterpret_cast<int16_t*>(&q);
#pragma unroll
            for (int k = 0; k < 8; k++)
                if (i0 + k < 0 || i0 + k >= (int)len) e[k] = 0;
        }
        uint32_t m2 = __vmaxs2(absmax0_2((uint32_t)q.x), absmax0_2((uint32_t)q.y));
        m2 = __vmaxs2(m2, __vmaxs2(absmax0_2((uint32_t)q.z), absmax0_2((uint32_t)q.w)));
        pk2 = __vmaxs2(pk2, m2);
        if (have_vm) vm[j] = (uint16_t)max(m2 & 0xffffu, m2 >> 16);
    }
    int pk = max((int)(pk2 & 0xffffu), (int)(pk2 >> 16));
    pk = block_allreduce<ASM_THREADS>(pk, OpMaxI32(), reinterpret_cast<int*>(sm.red));
    if (pk == 0) return len;
    const int limit = (int)f2s((float)pk * A.prm.silence_threshold);
    uint32_t keep_n = min_sil / 4;
    if (keep_n < 10) keep_n = 10;

    uint32_t* words;
    if (2 * wn <= SCR_WORDS) words = sm.scratch;
    else words = A.trim_scratch + (size_t)big * A.trim_scratch_words;   // host sized it for this task
    uint32_t* woff = words + wn;
    uint8_t* bytes = reinterpret_cast<uint8_t*>(words);
    const uint32_t nb = wn * 4;
    // mask byte t = 8 samples 8t .. 8t+7 of the region, 1 = |x| <= threshold: packed |x| (clamped at 0:
    // abs16(-32768) is negative, hence silent like 0), minus (limit + 1): the sign bit is the answer
    const uint32_t c2 = (uint32_t)((-(limit + 1)) & 0xffff) * 0x10001u;
    auto mask_byte = [&](uint32_t t) -> uint32_t {
        if (8 * t >= len) return 0u;
        int4 q = grid[t];
        if (phase) {
            int4 hi = make_int4(0, 0, 0, 0);
            if (8 * (t + 1) < len + phase) hi = grid[t + 1];
            q = shift_pick(q, hi, phase);
        }
        uint32_t acc = sign_mask2(__vadd2(absmax0_2((uint32_t)q.x), c2)) & 0x00020001u;
        acc |= sign_mask2(__vadd2(absmax0_2((uint32_t)q.y), c2)) & 0x00080004u;
        acc |= sign_mask2(__vadd2(absmax0_2((uint32_t)q.z), c2)) & 0x00200010u;
        acc |= sign_mask2(__vadd2(absmax0_2((uint32_t)q.w), c2)) & 0x00800040u;
        uint32_t byte = (acc | (acc >> 16)) & 0xffu;
        if (8 * t + 8 > len) byte &= (1u << (len - 8 * t)) - 1u;
        return byte;
    };

    // A run of >= min_sil silent samples covers >= needv = (min_sil - 7) / 8 whole grid vectors, all of
    // them silent, plus parts of the vector before and the vector after.  So the runs that can be cut are
    // found on the per-vector maxima (1 bit per vector), and the per-sample mask is only formed around
    // them: everywhere else it is left 0 ("not silent"), which can only split runs that are kept whole
    // anyway (ctts.c:1662-1668 copies every run shorter than min_sil).  No candidate: nothing is cut.
    constexpr uint32_t TRIM_MAX_CAND = 48;
    const uint32_t sv_words = (gvec + 31) >> 5;
    const uint32_t sv_at = 2 * wn + (gvec + 1) / 2 + 1;          // after the mask arrays and vm
    bool full_mask = true;
    if (have_vm && limit >= 0 && min_sil >= 15 + 8 * 14 && sv_at + sv_words + 2 + 2 * TRIM_MAX_CAND <= SCR_WORDS) {
        const uint32_t needv = (min_sil - 7) / 8;
        uint32_t* sv = sm.scratch + sv_at;
        uint32_t* cand = sv + sv_words;                           // [0] count, then (first, last) vector pairs
        for (uint32_t j0 = 0; j0 < gvec; j0 += ASM_THREADS) {
            const uint32_t j = j0 + tid;
            const bool sil = j < gvec && (int)vm[j] <= limit;
            const uint32_t w = __ballot_sync(0xffffffffu, sil);
            if (lane == 0 && (j0 >> 5) + warp < sv_words) sv[(j0 >> 5) + warp] = w;
        }
        if (tid == 0) cand[0] = 0;
        __syncthreads();
        int pushed = 0;
        for (uint32_t t = tid; t < sv_words; t += ASM_THREADS) {
            const uint32_t w = sv[t];
            uint32_t starts = w & ~((w << 1) | (t ? sv[t - 1] >> 31 : 0u));   // set bits whose predecessor is clear
            while (starts) {
                const uint32_t b = (uint32_t)__ffs(starts) - 1u;
                starts &= starts - 1u;
                uint32_t wi = t, sh = b, end;
                for (;;) {   // first clear bit at or after (wi, sh); bits past gvec are clear
                    const uint32_t z = ~sv[wi] & (0xffffffffu << sh);
                    if (z) { end = (wi << 5) + (uint32_t)__ffs(z) - 1u; break; }
                    if (++wi == sv_words) { end = sv_words << 5; break; }
                    sh = 0;
                }
                const uint32_t first = (t << 5) + b;
                if (end - first >= needv) {
                    pushed = 1;
                    const uint32_t k = atomicAdd(cand, 1u);
                    if (k < TRIM_MAX_CAND) {
                        cand[1 + 2 * k] = first;
                        cand[2 + 2 * k] = end - 1u;
                    }
                }
            }
        }
        // the decision to leave comes out of the barrier itself: a thread that returns goes on to reuse
        // the scratch, so nobody may still have to read the count from it
        if (!__syncthreads_or(pushed)) return len;
        const uint32_t n_cand = cand[0];
        if (n_cand <= TRIM_MAX_CAND) {
            full_mask = false;
            for (uint32_t i = tid; i < wn; i += ASM_THREADS) words[i] = 0u;
            __syncthreads();
            for (uint32_t k = 0; k < n_cand; k++) {
                // vectors first-1 .. last+1 hold region samples 8(first-1)-phase .. 8(last+2)-phase-1
                const uint32_t first = cand[1 + 2 * k], last = cand[2 + 2 * k];
                const uint32_t back = 1u + (phase ? 1u : 0u);
                const uint32_t t_lo = first > back ? first - back : 0u;
                const uint32_t t_hi = min(nb - 1u, last + 1u);
                for (uint32_t t = t_lo + tid; t <= t_hi; t += ASM_THREADS) bytes[t] = (uint8_t)mask_byte(t);
            }
        }
    }

    // 1 bit per sample: |x| <= threshold
    if (full_mask) {
        if (limit >= 0) {
            for (uint32_t t = tid; t < nb; t += ASM_THREADS) bytes[t] = (uint8_t)mask_byte(t);
        } else {
            for (uint32_t wd = warp; wd < wn; wd += ASM_THREADS / 32) {
                uint32_t i = (wd << 5) + lane;
                bool sil = (i < len) && (abs16(reg[i]) <= limit);
                uint32_t m = __ballot_sync(0xffffffffu, sil);
                if (lane == 0) words[wd] = m;
            }
        }
    }
    __syncthreads();

    // each thread owns a contiguous range of words
    const uint32_t per = (wn + ASM_THREADS - 1) / ASM_THREADS;
    const uint32_t j0 = min((uint32_t)tid * per, wn), j1 = min(j0 + per, wn);
    int my_last = -1, my_first = (int)len;
    for (uint32_t j = j0; j < j1; j++) {
        uint32_t valid = (j == wn - 1 && (len & 31)) ? ((1u << (len & 31)) - 1u) : 0xffffffffu;
        uint32_t ns = ~words[j] & valid;
        if (ns) {
            int l = (int)(j << 5) + 31 - __clz(ns);
            int f = (int)(j << 5) + __ffs(ns) - 1;
            my_last = l > my_last ? l : my_last;
            my_first = f < my_first ? f : my_first;
        }
    }
    int prev_ns = block_excl_scan<ASM_THREADS>(my_last, OpMaxI32(), -1, reinterpret_cast<int*>(sm.red), false);
    int next_ns = block_excl_scan<ASM_THREADS>(my_first, OpMinI32(), (int)len, reinterpret_cast<int*>(sm.red), true);

    // backward: first non-silent position after each owned word
    {
        int nx = next_ns;
        for (uint32_t j = j1; j > j0; j--) {
            uint32_t jj = j - 1;
            woff[jj] = (uint32_t)nx;
            uint32_t valid = (jj == wn - 1 && (len & 31)) ? ((1u << (len & 31)) - 1u) : 0xffffffffu;
            uint32_t ns = ~words[jj] & valid;
            if (ns) nx = (int)(jj << 5) + __ffs(ns) - 1;
        }
    }
    // forward: keep mask per word
    uint32_t kept = 0;
    {
        int pv = prev_ns;
        for (uint32_t j = j0; j < j1; j++) {
            uint32_t valid = (j == wn - 1 && (len & 31)) ? ((1u << (len & 31)) - 1u) : 0xffffffffu;
            uint32_t sil = words[j] & valid;
            uint32_t keep = ~sil & valid;
            int nx = (int)woff[j];
            uint32_t rem = sil;
            while (rem) {
                int lo = __ffs(rem) - 1;
                uint32_t t = ~(sil >> lo);
                int run_here = (t == 0u) ? (32 - lo) : (__ffs(t) - 1);
                int hi = lo + run_here;
                int start_g = (lo == 0) ? pv + 1 : (int)(j << 5) + lo;
                int end_g = (hi == 32) ? nx : (int)(j << 5) + hi;
                // the last word: a run touching the end of the region ends at len
                if (hi < 32 && (uint32_t)((j << 5) + hi) >= len) end_g = (int)len;
                uint32_t m_hi = (hi == 32) ? 0xffffffffu : ((1u << hi) - 1u);
                uint32_t m_lo = (1u << lo) - 1u;
                uint32_t run_mask = m_hi & ~m_lo;
                if ((uint32_t)(end_g - start_g) < min_sil) {
                    keep |= run_mask;
                } else {
                    int lim = start_g + (int)keep_n - (int)(j << 5);  // first bit NOT kept
                    if (lim > lo) {
                        int h2 = lim < hi ? lim : hi;
                        uint32_t m2 = (h2 >= 32) ? 0xffffffffu : ((1u << h2) - 1u);
                        keep |= m2 & ~m_lo;
                    }
                }
                rem &= ~run_mask;
            }
            uint32_t ns = ~sil & valid;
            if (ns) pv = (int)(j << 5) + 31 - __clz(ns);
            words[j] = keep;
            kept += __popc(keep);
        }
    }
    uint32_t total = 0;
    uint32_t off = block_excl_scan<ASM_THREADS>(kept, OpAddU32(), 0u, reinterpret_cast<uint32_t*>(sm.red), false, &total);
    for (uint32_t j = j0; j < j1; j++) {
        woff[j] = off;
        off += __popc(words[j]);
    }
    __syncthreads();
    if (total == len) return len;

    // in-place compaction: destinations never pass their sources, so chunks can
    // be processed in order with one barrier between a chunk's reads and writes
    for (uint32_t c0 = 0; c0 < len; c0 += ASM_THREADS * 8) {
        {
            // nothing dropped before the end of this chunk: every sample of it already is where it belongs
            const uint32_t e = c0 + ASM_THREADS * 8;
            if (e < len && woff[e >> 5] == e) continue;   // (e is a multiple of 32: woff is the kept count before it)
        }
        uint32_t i0 = c0 + (uint32_t)tid * 8;
        __align__(16) int16_t v[8];
        uint32_t km = 0, d0 = 0;
        if (i0 < len) {
            uint32_t j = i0 >> 5, b = i0 & 31;  // 8 | 32: one word
            uint32_t kw = words[j];
            km = (kw >> b) & 0xffu;
            d0 = woff[j] + __popc(kw & ((1u << b) - 1u));
            // the 8 samples at i0 (a multiple of 8) from the 16-byte grid of the region
            int4 q = grid[i0 >> 3];
            if (phase) {
                int4 hi = make_int4(0, 0, 0, 0);
                if (i0 + 8 < len + phase) hi = grid[(i0 >> 3) + 1];
                q = shift_pick(q, hi, phase);
            }
            *reinterpret_cast<int4*>(v) = q;
        }
        __syncthreads();
        if (km) {
#pragma unroll
            for (int k = 0; k < 8; k++)
                if (km & (1u << k)) reg[d0++] = v[k];
        }
    }
    __syncthreads();
    return total;
}

// apply_smooth_pitch_contour, ctts.c:2206, in gather form: every output sample
// collects the (at most two) 256-sample frames that cover it, in frame order;
// the int16 overlap-add wraps exactly as the reference's `+=` does.  In place,
// tile by tile: the originals a tile needs ([t0-256, t1+288)) are staged as floats
// in shared scratch (zero past the end of the segment: reads the reference performs
// past the end of its heap copy -- undefined behaviour there, DESIGN.md
// "Reference UB" -- yield 0 here), per-frame pitch factors come from a table.
// A thread's outputs are 256 apart, so its position inside a frame, the two window
// values and the norm (and its reciprocal for the division) are loop invariants.
// When `energy` is set the linear energy ramp of apply_phrase_intonation
// (ctts.c:2857-2864) over the whole word (index ebase + j, denominator eden) is
// applied to each sample as it is written.  Returns false if nothing was done.
constexpr uint32_t CONTOUR_TILE = ASM_THREADS * CONTOUR_KPT;
constexpr uint32_t CONTOUR_AHEAD = 528;                      // staged past a tile: frame reads reach i * pf <= 255 * 2.05
constexpr uint32_t CONTOUR_CARRY = PITCH_FRAME + 8 + CONTOUR_AHEAD;   // floats shared by consecutive tiles
constexpr uint32_t CONTOUR_STAGE = CONTOUR_TILE + CONTOUR_CARRY;  // 256 behind, 8 phase, tile, ahead
static_assert(CONTOUR_CARRY % 8 == 0, "carry is whole staging vectors");
constexpr uint32_t CONTOUR_PF_MAX = 1024;                    // frames with a tabulated pitch factor
constexpr uint32_t CONTOUR_SCRATCH_WORDS = CONTOUR_STAGE + CONTOUR_PF_MAX;
// the unit-head staging area follows the scratch and is dead during WORD_END: at least 252 more words (hcap >= 504 samples)
static_assert(CONTOUR_SCRATCH_WORDS + 8 <= SCR_WORDS + 252, "contour scratch fits");
static_assert(ASM_THREADS % 128 == 0, "a thread's outputs must keep their position inside a frame");

// one frame's contribution to an output sample (ctts.c:2236-2251): linear-interpolated read at
// i * pf inside the frame that starts at fb[0], windowed, truncated to int
__device__ __forceinline__ int contour_term(const float* fb, float fi, float pf, float w) {
    const float xs = fi * pf;
    const int k = (int)xs;
    const float fr = xs - (float)k;
    const float s0 = fb[k], s1 = fb[k + 1];
    const float v = (k + 1 < PITCH_FRAME) ? s0 * (1.0f - fr) + s1 * fr : s0;
    return (int)(v * w);
}

__device__ bool pitch_contour(const Smem& sm, int16_t* x, uint32_t n, float f0, float f1, bool energy,
                              float e0, float de, float eden, uint32_t ebase) {
    if (n < 100 || fabsf(f0 - f1) < 0.01f) return false;
    if (n < PITCH_FRAME) return false;  // no frame fits: every sample keeps its original value
    const int tid = threadIdx.x;
    const uint32_t frames = (n - PITCH_FRAME) / (PITCH_FRAME / 2) + 1;
    const bool degenerate = (n == PITCH_FRAME);  // 1/(n-256) = inf in the reference: NaN indices
    const float inv = 1.0f / (float)(n - PITCH_FRAME);
    float* stage = reinterpret_cast<float*>(sm.scratch);
    float* pft = stage + CONTOUR_STAGE;
    const bool tabulated = frames <= CONTOUR_PF_MAX;
    if (tabulated) {
        for (uint32_t k = tid; k < frames; k += ASM_THREADS) {
            float t = (float)(k << 7) * inv;
            float st = t * t * (3.0f - 2.0f * t);
            pft[k] = f0 + (f1 - f0) * st;
        }
    }
    // loop invariants of this thread
    const uint32_t i1 = (uint32_t)tid & 127u;
    const float w_lo = sm.hann256[i1], w_hi = sm.hann256[i1 + 128], w_2 = sm.nrm2[i1];
    const float r_lo = recip_for_div(w_lo), r_hi = recip_for_div(w_hi), r_2 = recip_for_div(w_2);
    const float fi_lo = (float)i1, fi_hi = (float)(i1 + 128u);
    const float r_e = recip_for_div(eden);
    // stage[phase + 256 + u] = x[t0 + u]: aligned 8-sample vectors of x land on float4 pairs
    const uint32_t phase = (uint32_t)((reinterpret_cast<uintptr_t>(x) >> 1) & 7u);
    float* sbase = stage + phase + PITCH_FRAME;   // index u relative to the tile start
    static_assert(CONTOUR_CARRY / 4 <= ASM_THREADS, "one float4 of carry per thread");
    float4 carry = make_float4(0.f, 0.f, 0.f, 0.f);
    for (uint32_t t0 = 0; t0 < n; t0 += CONTOUR_TILE) {
        const uint32_t t1 = min(t0 + CONTOUR_TILE, n);
        const bool more = t0 + CONTOUR_TILE < n;
        // ---- staging: what overlaps the previous tile was picked up in a register before that
        //      tile's closing barrier (no barrier of its own), the rest is loaded and converted
        if (t0 != 0 && tid < (int)(CONTOUR_CARRY / 4)) *(reinterpret_cast<float4*>(stage) + tid) = carry;
        for (uint32_t v = (t0 == 0 ? 0u : CONTOUR_CARRY / 8) + tid; v < CONTOUR_STAGE / 8; v += ASM_THREADS) {
            const int u0 = (int)(v << 3) - (int)(phase + PITCH_FRAME);   // u of the vector's first sample
            const long long g0 = (long long)t0 + u0;                     // segment index
            int4 q = make_int4(0, 0, 0, 0);
            if (g0 >= 0 && g0 + 8 <= (long long)n) {
                q = *reinterpret_cast<const int4*>(x + g0);
            } else if (g0 + 8 > 0 && g0 < (long long)n) {
                int16_t* e = reinterpret_cast<int16_t*>(&q);
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    long long g = g0 + k;
                    if (g >= 0 && g < (long long)n) e[k] = x[g];
                }
            }
            float4 a, b;
            a.x = (float)(short)((uint32_t)q.x & 0xffffu); a.y = (float)(short)((uint32_t)q.x >> 16);
            a.z = (float)(short)((uint32_t)q.y & 0xffffu); a.w = (float)(short)((uint32_t)q.y >> 16);
            b.x = (float)(short)((uint32_t)q.z & 0xffffu); b.y = (float)(short)((uint32_t)q.z >> 16);
            b.z = (float)(short)((uint32_t)q.w & 0xffffu); b.w = (float)(short)((uint32_t)q.w >> 16);
            *(reinterpret_cast<float4*>(stage) + 2 * v) = a;
            *(reinterpret_cast<float4*>(stage) + 2 * v + 1) = b;
        }
        __syncthreads();
        // ---- outputs of this tile
        // A 256-output block is "interior" when each of its outputs is covered by two whole
        // frames: branch-free code.  In an interior tile the four outputs of a thread are
        // independent instruction streams.
        const uint32_t kb = (t0 >> 7) + ((uint32_t)tid >> 7);           // frame k1 of output r = 0
        const float* fb0 = sbase + (int)((uint32_t)tid & ~127u) - 128;  // start of frame k1 - 1, r = 0
        const bool can_fast = tabulated && !degenerate && w_2 > 0.01f;
        const bool fast = can_fast && t0 >= 128 && t0 + CONTOUR_TILE <= n && ((t0 + CONTOUR_TILE - 1) >> 7) < frames;
        if (fast) {
            int o[CONTOUR_KPT];
#pragma unroll
            for (int r = 0; r < CONTOUR_KPT; r++) {
                const float pfa = pft[kb + 2 * r - 1], pfb = pft[kb + 2 * r];
                const float* fa = fb0 + ASM_THREADS * r;
                const int ta = contour_term(fa, fi_hi, pfa, w_hi);
                const int tb = contour_term(fa + 128, fi_lo, pfb, w_lo);
                const int acc = (int)(int16_t)(ta + tb);
                o[r] = cvt_sat_s16(div_by((float)acc, w_2, r_2));
            }
            if (energy) {
#pragma unroll
                for (int r = 0; r < CONTOUR_KPT; r++) {
                    const float t = div_by((float)(t0 + (uint32_t)tid + (uint32_t)(ASM_THREADS * r) + ebase), eden, r_e);
                    o[r] = cvt_sat_s16((float)o[r] * (e0 + de * t));
                }
            }
#pragma unroll
            for (int r = 0; r < CONTOUR_KPT; r++) x[t0 + (uint32_t)tid + (uint32_t)(ASM_THREADS * r)] = (int16_t)o[r];
            if (more && tid < (int)(CONTOUR_CARRY / 4)) carry = *(reinterpret_cast<const float4*>(stage + CONTOUR_TILE) + tid);
            __syncthreads();
            continue;
        }
#pragma unroll 1
        for (int r = 0; r < CONTOUR_KPT; r++) {
            const uint32_t j0 = t0 + (uint32_t)r * ASM_THREADS;
            if (can_fast && j0 >= 128 && j0 + ASM_THREADS <= n && ((j0 + ASM_THREADS - 1) >> 7) < frames) {
                // interior block of an edge tile
                const float pfa = pft[kb + 2 * r - 1], pfb = pft[kb + 2 * r];
                const float* fa = fb0 + ASM_THREADS * r;
                const int ta = contour_term(fa, fi_hi, pfa, w_hi);
                const int tb = contour_term(fa + 128, fi_lo, pfb, w_lo);
                int o = cvt_sat_s16(div_by((float)(int)(int16_t)(ta + tb), w_2, r_2));
                if (energy) {
                    const float t = div_by((float)(j0 + (uint32_t)tid + ebase), eden, r_e);
                    o = cvt_sat_s16((float)o * (e0 + de * t));
                }
                x[j0 + (uint32_t)tid] = (int16_t)o;
                continue;
            }
            const uint32_t ju = (uint32_t)tid + (uint32_t)r * ASM_THREADS;   // index inside the tile
            const uint32_t j = t0 + ju;
            if (j >= t1) continue;
            const uint32_t k1 = j >> 7;
            const bool vb = k1 < frames;                    // frame k1 covers j at i = i1
            const bool va = k1 >= 1 && k1 - 1 < frames;     // frame k1-1 covers j at i = i1 + 128
            int acc = 0;
            float norm = 0.0f, rn = 0.0f;
            if (va) {
                float pf;
                if (tabulated) pf = pft[k1 - 1];
                else {
                    float t = (float)((k1 - 1) << 7) * inv;
                    pf = f0 + (f1 - f0) * (t * t * (3.0f - 2.0f * t));
                }
                if (!degenerate) acc = contour_term(sbase + ((int)ju - (int)i1 - 128), fi_hi, pf, w_hi);
                norm = w_hi;
                rn = r_hi;
            }
            if (vb) {
                float pf;
                if (tabulated) pf = pft[k1];
                else {
                    float t = (float)(k1 << 7) * inv;
                    pf = f0 + (f1 - f0) * (t * t * (3.0f - 2.0f * t));
                }
                int tb = 0;
                if (!degenerate) tb = contour_term(sbase + ((int)ju - (int)i1), fi_lo, pf, w_lo);
                acc = (int)(int16_t)((int16_t)acc + (int16_t)tb);
                norm = va ? w_2 : w_lo;
                rn = va ? r_2 : r_lo;
            } else {
                acc = (int)(int16_t)acc;
            }
            int o;
            if (norm > 0.01f) o = cvt_sat_s16(div_by((float)acc, norm, rn));
            else o = (int)sbase[ju];
            if (energy) {
                const float t = div_by((float)(j + ebase), eden, r_e);
                o = cvt_sat_s16((float)o * (e0 + de * t));
            }
            x[j] = (int16_t)o;
        }
        if (more && tid < (int)(CONTOUR_CARRY / 4)) carry = *(reinterpret_cast<const float4*>(stage + CONTOUR_TILE) + tid);
        __syncthreads();
    }
    return true;
}

// ctts.c:3693-3713 / :3878-3898: trim then phrase intonation on [word_start, count).  The two halves can be
// run apart: a canonical region stops after the trim, a task that takes its samples from the region store
// resumes at the intonation.
__device__ void op_word_end(State& s, const Smem& sm, const AsmArgs& A, uint32_t big, const ctts_plan_op& op,
                            bool do_trim = true, bool do_inton = true) {
    const int tid = threadIdx.x;
    if (do_trim && (op.flags & CTTS_WE_TRIM) && s.cnt > s.word_start) {
        uint32_t len = s.cnt - s.word_start;
        if (len > A.prm.min_silence_samples)
            s.cnt = s.word_start + trim_region(sm, A, big, s.w + s.word_start, len);
    }
    if (!do_inton || s.cnt <= s.word_start) return;
    const uint32_t n = s.cnt - s.word_start;
    int16_t* x = s.w + s.word_start;
    // device half of apply_phrase_intonation, ctts.c:2740, :2774-2790, :2839-2865
    if (!(op.flags & CTTS_WE_INTON) || n < 100) return;
    const bool energy = (op.flags & CTTS_WE_ENERGY) != 0;
    const float e0 = op.e0, de = op.e1 - op.e0;
    const float den = (float)(n - 1);
    bool done = false;
    // [lo, hi): samples whose energy ramp is still to be applied after the contour
    uint32_t lo = 0, hi = n;
    if (op.flags & CTTS_WE_CIRCUMFLEX) {
        uint32_t rise = (uint32_t)(unsigned long long)((float)n * 0.6f);
        if (rise > 100 && n - rise > 100) {
            bool a = pitch_contour(sm, x, rise, op.f0, op.f2, energy, e0, de, den, 0);
            bool b = pitch_contour(sm, x + rise, n - rise, op.f2, op.f1, energy, e0, de, den, rise);
            if (a) lo = rise;
            if (b) hi = rise;
            if (a && b) hi = lo = 0;
            if (!a && b) { lo = 0; hi = rise; }
            if (a && !b) { lo = rise; hi = n; }
            done = true;
        }
    }
    if (!done && pitch_contour(sm, x, n, op.f0, op.f1, energy, e0, de, den, 0)) lo = hi = 0;
    if (energy && hi > lo) {
        for (uint32_t i = lo + tid; i < hi; i += ASM_THREADS) {
            float t = (float)i / den;
            float e = e0 + de * t;
            x[i] = f2s(clamp16f((float)x[i] * e));
        }
        __syncthreads();
    }
}


}  // namespace ctts
