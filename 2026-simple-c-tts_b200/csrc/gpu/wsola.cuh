// wsola.cuh -- time stretching (time_stretch, ctts.c:3490-3617).
//
// The analysis position of frame k depends on where frame k-1 was taken (ctts.c:3555-3592), but the
// transition is a pure function T_k(offset of frame k-1), and on real signals it has a fixed point:
// the candidate at the previous frame's offset IS the target (same samples), so it scores exactly
// 1.0f (sum_prod == sum_sq1 == sum_sq2 bit for bit and sqrtf(s*s) == s), nothing can score higher
// than 1 except by rounding, and the offset repeats.  Digital silence behind the frame makes every
// score 0, the first candidate in bounds (-128) wins (ctts.c:3426, :3452-3461) and then sticks.  So
// the chain is SPECULATED and VERIFIED instead of walked:
//   wsola_scan_kernel    per utterance: the frame count and F = the first frame whose target, with
//                        every offset before it 0, is all zero.  Speculated offsets: 0 before F,
//                        -128 from F on.
//   wsola_verify_kernel  every frame independently (tiles of 24 frames per CTA): under the
//                        speculated previous offset h the frame's offset is h again iff every other
//                        candidate of the coarse and the fine stage scores < 1.0f.  Tier 1 proves that
//                        with a partial sum: 1 - corr = |x^ - t^|^2 / 2 >= sum over three 12-term blocks
//                        of (x^_i - t^_i)^2 (x^, t^ the unit-normalised windows; window energies exact,
//                        from a 64-bit prefix sum of squares): 36 of the 384 terms -- and, the blocks
//                        being one analysis hop apart, a partial dot product serves three frames.  What
//                        tier 1 cannot reject (0.45 candidates per frame) gets the full 384-term FMA
//                        filter score (tier 2, one warp per candidate), and what that leaves within eps
//                        of 1 the reference's exact loop (tier 3).  A candidate that really reaches 1.0f
//                        marks the frame as the utterance's first bad frame.
//   wsola_search_kernel  the repair: the frame-by-frame chain walk (below), started at the first bad
//                        frame with the verified position before it; exits at once when the
//                        speculation held (every utterance of the benchmark workloads).
//
//   wsola_ola_kernel     overlap-add in gather form (below).
//
// wsola_search_kernel: one CTA per stretched utterance walks the frames in order.  The reference
//    scores <= 65 coarse candidates (offsets -128..128 step 4) and then <= 6 fine ones around the best, each score a 384-term sequential float loop
//    (groups of 4: ((p0+p1)+p2)+p3, then +=, ctts.c:3411-3413) -- 91 % of its run time.
//    Here every frame is decided in two steps that give the identical result:
//      FILTER  all candidates get an approximate score: the cross term by FMA in any order
//              (4 candidates per thread share the operand loads, the 384 terms are split
//              over 8 threads), the two energies EXACTLY from a 64-bit prefix sum of squares
//              of the 640 samples the frame can see.  |approx - reference| <= 2e-5
//              (summation error bound as in asm_pitch.cuh; eps = 1e-4 is used).
//      DECIDE  candidates whose interval [a-eps, a+eps] reaches the best lower bound are the
//              only possible winners.  One candidate: done.  Several: those (typically 2-3)
//              are evaluated with the reference's exact loop, one thread each, and compared
//              exactly (largest score, first in scan order on ties).  A score whose
//              denominator is clearly < 1 is exactly 0 (ctts.c:3426) and needs no evaluation.
//    Consecutive views overlap by 512 samples: the samples (as floats) and the prefix sums live
//    in rings, a frame appends only its 128 new samples, prefetched while the previous frame is
//    decided.  Output: the analysis position of every frame.
// wsola_ola_kernel: embarrassingly parallel gather-form overlap-add.  Each
//    output sample adds its <= 8 windowed frame contributions in frame order
//    into an int16 accumulator that wraps exactly like the reference's `+=`
//    (ctts.c:3577) and a float norm (ctts.c:3578), normalises, and the CTA
//    reports the last non-zero sample for the trailing-zero trim (ctts.c:3611).
#pragma once
#include <climits>
#include <cstdint>
#include <cuda_runtime.h>

#include "asm_common.cuh"

namespace ctts {

constexpr int WS_FRAME = 512;    // ctts.c:3506
constexpr int WS_HOP = 128;      // analysis hop
constexpr int WS_OVERLAP = 384;  // correlation length
constexpr int WS_SHIFT = 128;    // +-search range
constexpr int WS_RANGE = 2 * WS_SHIFT + WS_OVERLAP;  // 640 samples visible to the candidates
constexpr int WS_CTAS_PER_SM = 10;   // resident search CTAs per SM (registers and shared memory allow 10)
constexpr int WS_THREADS = 128;  // 16 candidate groups x 8 splits; the 65th candidate is spread over all threads
constexpr int WS_GROUPS = 16;    // groups of 4 coarse candidates (offsets -128 .. 124)
constexpr int WS_SPLITS = 8;     // threads sharing the 384 terms of a group
constexpr int WS_FINE_SPLITS = 16;
constexpr int WS_RING = 1024;    // samples kept: a double-mapped ring makes every 640-sample view contiguous
constexpr int WS_MAXC = 72;      // candidates of one decision
constexpr float WS_EPS = 1e-4f;
constexpr int OLA_THREADS = 256;
constexpr int OLA_SPT = 16; // rows (output samples) per thread
constexpr int OLA_QMAX = 8;  // frames that can cover one output sample: 512 / 64 (hop >= 64 on the tiled path)

struct StretchTask {
    uint32_t utt;
    uint32_t hop;                 // synthesis hop = (size_t)(128 / speed)
    unsigned long long pre_off;   // pre-stretch buffer offset in dst_pre
    unsigned long long out_off;   // output slot offset in dst_final
    uint32_t out_cap;
    uint32_t pos_off;             // offset into frame_pos
    uint32_t max_frames;
};

struct WsolaArgs {
    const StretchTask* tasks;
    uint32_t n_tasks;      // of the whole plan
    uint32_t task_first;   // search launch: CTA b works on task task_first + b
    const int16_t* pre;
    const uint32_t* pre_counts;
    int16_t* out;
    uint32_t* out_counts;
    uint32_t* frame_pos;
    uint32_t* n_frames;   // per task; [n_tasks + task] = decisions that needed an exact evaluation
    uint32_t* first_bad;    // per task: first frame the speculation could not confirm (0xffffffff: none)
    uint32_t* first_silent; // per task: F of wsola_scan_kernel (0xffffffff: none)
    uint32_t* tier2;        // per task: candidates tier 1 could not reject
    uint32_t task_count;    // tasks of this launch
    uint32_t speculate;     // 0: every utterance is walked by the chain kernel from frame 1 (debug / tests)
    uint32_t force_bad;     // != 0: every frame k with k % force_bad == 0 is reported bad (tests of the repair path)
    const float* hann512;
    const uint32_t* ola_block_task;   // per OLA block of this launch: task index (plan-wide)
    const uint32_t* ola_block_first;  // per OLA block: first output sample
};

// cross_correlation (ctts.c:3390) sums of one candidate in the reference's order.
// a = candidate window, b = target; when b == nullptr only the energy of a is formed.
__device__ __forceinline__ void ws_exact_sums(const float* a, const float* b, float* sp_out, float* sa_out) {
    float sp = 0.0f, sa = 0.0f;
#pragma unroll 2
    for (int m = 0; m < WS_OVERLAP / 4; m++) {
        const float a0 = a[4 * m], a1 = a[4 * m + 1], a2 = a[4 * m + 2], a3 = a[4 * m + 3];
        sa += a0 * a0 + a1 * a1 + a2 * a2 + a3 * a3;
        if (b) {
            const float b0 = b[4 * m], b1 = b[4 * m + 1], b2 = b[4 * m + 2], b3 = b[4 * m + 3];
            sp += a0 * b0 + a1 * b1 + a2 * b2 + a3 * b3;
        }
    }
    *sp_out = sp;
    *sa_out = sa;
}

// One decision of find_best_match_wsola (ctts.c:3436): `cnt` candidates in the reference's scan
// order, each with an approximate score ca[i] and an error radius ce[i] (0: the score is exact).
// The winner is the first candidate in scan order that holds the maximum exact score.  Returns
// its index; CTA-uniform; all threads must call.
struct WsDecide {
    float* ca;       // [WS_MAXC]
    float* ce;       // [WS_MAXC]
    int* xoff;       // [WS_MAXC] start of the candidate's window in xin
    int* list;       // [WS_MAXC]
    int* cnt_s;      // [4] shared: survivors, inexact survivors, fast-path winner, lower bound
    float* sb_s;     // exact target energy (valid after the first exact evaluation of a frame)
    int* sb_valid;
    uint32_t* exact_count;
};

// order-preserving map float -> uint32
__device__ __forceinline__ uint32_t ws_sortable(float v) {
    const uint32_t b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

__device__ int ws_decide(const WsDecide& D, int cnt, const float* xin, const float* tgt) {
    const int tid = threadIdx.x;
    // ---- warp 0, three candidates per lane: the best lower bound, who survives it, and -- if every
    //      survivor's score is already exact, or there is only one -- the winner
    if (tid < 32) {
        float a[3], e[3];
        uint32_t lbk = 0;   // sortable key of the best lower bound
#pragma unroll
        for (int r = 0; r < 3; r++) {
            const int i = tid + 32 * r;
            a[r] = i < cnt ? D.ca[i] : -3.0f;
            e[r] = i < cnt ? D.ce[i] : 0.0f;
            lbk = max(lbk, ws_sortable(a[r] - e[r]));
        }
        lbk = __reduce_max_sync(0xffffffffu, lbk);
        int ns = 0, inexact = 0;
        uint32_t best_k = 0;      // best score among this lane's survivors
        uint32_t best_i = 0xffffu;
#pragma unroll
        for (int r = 0; r < 3; r++) {
            const int i = tid + 32 * r;
            const bool in = i < cnt && ws_sortable(a[r] + e[r]) >= lbk;
            ns += __popc(__ballot_sync(0xffffffffu, in));
            inexact += __popc(__ballot_sync(0xffffffffu, in && e[r] != 0.0f));
            const uint32_t sk = ws_sortable(a[r]);
            if (in && sk > best_k) {   // ascending i inside a lane: strict > keeps the first
                best_k = sk;
                best_i = (uint32_t)i;
            }
        }
        // (score, scan position): the best score, first in scan order among equals
        const uint32_t top = __reduce_max_sync(0xffffffffu, best_k);
        const uint32_t first = __reduce_min_sync(0xffffffffu, best_k == top ? best_i : 0xffffu);
        if (tid == 0) {
            D.cnt_s[0] = ns;
            D.cnt_s[1] = inexact;
            D.cnt_s[2] = (int)first;
            D.cnt_s[3] = (int)lbk;
        }
    }
    __syncthreads();
    const int ns = D.cnt_s[0], inexact = D.cnt_s[1];
    if (ns == 1 || inexact == 0) {
        const int w = D.cnt_s[2];
        __syncthreads();
        return w;
    }
    // ---- several possible winners, not all exact: the reference's loop for those, one thread each
    const uint32_t lbk = (uint32_t)D.cnt_s[3];
    if (tid < 32) {   // survivors in scan order (a warp-ordered compaction keeps the order)
        int base = 0;
        for (int i0 = 0; i0 < cnt; i0 += 32) {
            const int i = i0 + tid;
            const bool in = i < cnt && ws_sortable(D.ca[i] + D.ce[i]) >= lbk;
            const uint32_t m = __ballot_sync(0xffffffffu, in);
            if (in) D.list[base + __popc(m & ((1u << tid) - 1u))] = i;
            base += __popc(m);
        }
    }
    if (tid == 0) atomicAdd(D.exact_count, 1u);
    const bool need_sb = *D.sb_valid == 0;
    __syncthreads();
    // job 0: the target's energy (once per frame); jobs 1..ns: the survivors
    for (int j = tid; j <= ns; j += WS_THREADS) {
        if (j == 0) {
            if (need_sb) {
                float sp, sb;
                ws_exact_sums(tgt, nullptr, &sp, &sb);
                *D.sb_s = sb;
                *D.sb_valid = 1;
            }
        } else {
            const int i = D.list[j - 1];
            if (D.ce[i] != 0.0f) {
                float sp, sa;
                ws_exact_sums(xin + D.xoff[i], tgt, &sp, &sa);
                D.ca[i] = sp;             // parked: the score needs sb
                D.ce[i] = -sa - 1.0f;     // < 0 marks "raw sums"
            }
        }
    }
    __syncthreads();
    for (int j = tid; j < ns; j += WS_THREADS) {
        const int i = D.list[j];
        if (D.ce[i] < 0.0f) {
            const float sa = -(D.ce[i] + 1.0f);
            const float den = sqrtf(sa * *D.sb_s);
            D.ca[i] = den < 1.0f ? 0.0f : D.ca[i] / den;
            D.ce[i] = 0.0f;
        }
    }
    __syncthreads();
    // first in scan order with the maximum exact score
    int w = D.list[0];
    float best = D.ca[w];
    for (int j = 1; j < ns; j++) {
        const int i = D.list[j];
        if (D.ca[i] > best) {
            best = D.ca[i];
            w = i;
        }
    }
    __syncthreads();
    return w;
}

// exact integer (< 2^40) -> float, two conversions instead of the 64-bit library path; the two
// roundings (<= 2^-24 each) are inside the filter's error budget
__device__ __forceinline__ float ws_u64_to_float(unsigned long long v) {
    return __fmaf_rn((float)(uint32_t)(v >> 32), 4294967296.0f, (float)(uint32_t)v);
}

__global__ void __launch_bounds__(WS_THREADS, WS_CTAS_PER_SM) wsola_search_kernel(const WsolaArgs A) {
    // ring of the last 1024 input samples as floats, stored twice (i and i + 1024): the view of a
    // frame, input[nominal-128 .. nominal+512), is 640 contiguous floats at a 16-byte aligned offset
    __shared__ __align__(16) float xr[2 * WS_RING];
    // Pr[p & 1023] = sum_{j<p} input[j]^2 (exact, 64-bit): every window energy is a difference
    __shared__ __align__(16) unsigned long long Pr[WS_RING];
    __shared__ __align__(16) float tgt[WS_OVERLAP];    // previous frame's last 384 samples (aligned copy)
    __shared__ unsigned long long s_wtot[WS_THREADS / 32];
    __shared__ unsigned long long s_ptotal;
    __shared__ float ca[WS_MAXC], ce[WS_MAXC];
    __shared__ int xoff[WS_MAXC], list[WS_MAXC], coff[WS_MAXC];
    __shared__ float s_c64[WS_GROUPS];                 // partial cross terms of the 65th candidate
    __shared__ int s_cnt[4], s_sb_valid;
    __shared__ float s_sb;

    const uint32_t ti = A.task_first + blockIdx.x;
    const StretchTask task = A.tasks[ti];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t n = A.pre_counts[task.utt];
    const int16_t* in = A.pre + task.pre_off;
    uint32_t* fpos = A.frame_pos + task.pos_off;
    uint32_t* exact_count = A.n_frames + A.n_tasks + ti;

    uint32_t frames = n >= WS_FRAME ? (n - WS_FRAME) / WS_HOP + 1 : 0;
    if (frames > task.max_frames) frames = task.max_frames;
    // the chain is walked from the first frame the speculation could not confirm (wsola_verify_kernel);
    // every position before it is verified
    const uint32_t k0 = A.first_bad[ti];
    if (k0 >= frames) return;   // (k0 >= 1; CTA-uniform)
    const uint32_t vstart0 = (k0 - 1u) * WS_HOP;   // the view of frame k0 starts here
    if (tid == 0) {
        s_ptotal = 0ull;
        Pr[vstart0 & (WS_RING - 1)] = 0ull;
    }
    uint32_t prev_pos = fpos[k0 - 1u];   // written by an earlier kernel

    WsDecide D{ca, ce, xoff, list, s_cnt, &s_sb, &s_sb_valid, exact_count};

    // append 128 samples (one per thread) at absolute positions base .. base+127 to both rings.
    // Two barriers inside; the caller synchronises before anything reads the rings.
    auto append = [&](uint32_t base, float v) {
        const uint32_t p = base + (uint32_t)tid;
        xr[p & (WS_RING - 1)] = v;
        xr[(p & (WS_RING - 1)) + WS_RING] = v;
        const int iv = (int)v;
        const unsigned long long q = (unsigned long long)(uint32_t)(iv * iv);
        unsigned long long inc = q;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) s_wtot[warp] = inc;
        __syncthreads();
        unsigned long long carry = s_ptotal;
#pragma unroll
        for (int w = 0; w < WS_THREADS / 32; w++)
            if (w < warp) carry += s_wtot[w];
        Pr[(p + 1) & (WS_RING - 1)] = carry + inc;
        __syncthreads();
        if (tid == WS_THREADS - 1) s_ptotal = carry + inc;
    };
    auto load = [&](uint32_t base) {
        const uint32_t p = base + (uint32_t)tid;
        return p < n ? (float)in[p] : 0.0f;
    };
    // frame k0 sees input[vstart0 .. vstart0 + 640) (the prefix sums start at its first sample: only
    // differences are used)
    for (uint32_t b = 0; b < WS_RANGE; b += WS_THREADS) {
        append(vstart0 + b, load(vstart0 + b));
        __syncthreads();
    }
    // frame k > k0 adds input[128 k + 384 .. 128 k + 512); loaded two frames ahead (the buffer is in HBM)
    float next_v = frames > k0 + 1 ? load(vstart0 + WS_RANGE) : 0.0f;
    float next_v2 = frames > k0 + 2 ? load(vstart0 + WS_RANGE + WS_THREADS) : 0.0f;

    for (uint32_t k = k0; k < frames; k++) {
        const int nominal = (int)(k * WS_HOP);
        const uint32_t vstart = (uint32_t)(nominal - WS_SHIFT);           // absolute position of the view
        if (k > k0) {
            append(vstart + WS_RANGE - WS_THREADS, next_v);              // the 128 samples this view adds
            next_v = next_v2;
            if (k + 2 < frames) next_v2 = load(vstart + WS_RANGE + WS_THREADS);   // two frames ahead
        }
        const float* xin = xr + (vstart & (WS_RING - 1));
        if (tid == 0) s_sb_valid = 0;
        __syncthreads();
        // the target inside the view: in[prev_pos + 128 ..) = xin[ti ..], ti in [0, 256]
        const int ti = (int)prev_pos + (WS_FRAME - WS_OVERLAP) - (nominal - WS_SHIFT);
        int tgt_nz = 0;
        for (int i = tid; i < WS_OVERLAP; i += WS_THREADS) {
            const float v = xin[ti + i];
            tgt[i] = v;
            tgt_nz |= v != 0.0f;
        }
        // scan order of the coarse stage: ascending offset, candidates out of bounds skipped (ctts.c:3452)
        // candidate c (offset -128 + 4c) is in bounds iff cpos >= 0 && cpos + 512 <= n
        const int c_first = 0;                                             // nominal >= 128 for k >= 1
        int c_last = 2 * WS_SHIFT / 4;
        const long long room = (long long)n - WS_FRAME - nominal;   // largest offset in bounds (>= 0)
        if (room < WS_SHIFT) c_last = (int)((room + WS_SHIFT) / 4);
        if (!__syncthreads_or(tgt_nz)) {
            // digital silence behind us: every denominator is 0 < 1, every score is exactly 0
            // (ctts.c:3426), the first candidate in bounds wins and no fine candidate beats it
            uint32_t pos = (uint32_t)(nominal - WS_SHIFT + 4 * c_first);
            if (tid == 0) fpos[k] = pos;
            prev_pos = pos;
            continue;
        }

        // ---- FILTER, coarse: group g = candidates 4g..4g+3 (window starts 16g + 4j), split s
        float sp[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        const int g = tid >> 3, sidx = tid & 7;
        {
            // split s takes terms [52 s, 52 s + 52) (the last one 20): starts 13 float4 apart put the
            // eight lanes of a quarter warp on eight different bank groups (12 would collide 4-way)
            const int m_cnt = sidx == WS_SPLITS - 1 ? WS_OVERLAP / 4 - 13 * (WS_SPLITS - 1) : 13;
            const float4* x4 = reinterpret_cast<const float4*>(tgt) + 13 * sidx;
            const float4* y4 = reinterpret_cast<const float4*>(xin) + 4 * g + 13 * sidx;
            float4 y0 = y4[0], y1 = y4[1], y2 = y4[2];
            for (int m = 0; m < m_cnt; m++) {
                const float4 x = x4[m];
                const float4 y3 = y4[m + 3];
                sp[0] = __fmaf_rn(x.x, y0.x, sp[0]); sp[0] = __fmaf_rn(x.y, y0.y, sp[0]);
                sp[0] = __fmaf_rn(x.z, y0.z, sp[0]); sp[0] = __fmaf_rn(x.w, y0.w, sp[0]);
                sp[1] = __fmaf_rn(x.x, y1.x, sp[1]); sp[1] = __fmaf_rn(x.y, y1.y, sp[1]);
                sp[1] = __fmaf_rn(x.z, y1.z, sp[1]); sp[1] = __fmaf_rn(x.w, y1.w, sp[1]);
                sp[2] = __fmaf_rn(x.x, y2.x, sp[2]); sp[2] = __fmaf_rn(x.y, y2.y, sp[2]);
                sp[2] = __fmaf_rn(x.z, y2.z, sp[2]); sp[2] = __fmaf_rn(x.w, y2.w, sp[2]);
                sp[3] = __fmaf_rn(x.x, y3.x, sp[3]); sp[3] = __fmaf_rn(x.y, y3.y, sp[3]);
                sp[3] = __fmaf_rn(x.z, y3.z, sp[3]); sp[3] = __fmaf_rn(x.w, y3.w, sp[3]);
                y0 = y1; y1 = y2; y2 = y3;
            }
        }
        // the 65th candidate (offset +128, window xin[256 ..)): three terms per thread
        float s64 = 0.0f;
#pragma unroll
        for (int q = 0; q < WS_OVERLAP / WS_THREADS; q++)
            s64 = __fmaf_rn(tgt[tid + WS_THREADS * q], xin[2 * WS_SHIFT + tid + WS_THREADS * q], s64);
        // reduce the 8 splits of a group (adjacent lanes)
#pragma unroll
        for (int o = 1; o < WS_SPLITS; o <<= 1) {
#pragma unroll
            for (int j = 0; j < 4; j++) sp[j] += __shfl_xor_sync(0xffffffffu, sp[j], o);
            s64 += __shfl_xor_sync(0xffffffffu, s64, o);
        }
        if (sidx == 0) s_c64[g] = s64;
        const unsigned long long* Pv = Pr;   // window energy of view index a: P[vstart + a + 384] - P[vstart + a]
        auto energy = [&](int a) {
            const uint32_t p0 = vstart + (uint32_t)a;
            return ws_u64_to_float(Pv[(p0 + WS_OVERLAP) & (WS_RING - 1)] - Pv[p0 & (WS_RING - 1)]);
        };
        const float sb = energy(ti);
        __syncthreads();   // s_c64
        {
            // lanes 0..3 of a group score its four candidates; lane 4 of group 0 scores the 65th
            int c = -1;
            float spc = 0.0f;
            if (sidx < 4) {
                c = 4 * g + sidx;
                spc = sidx == 0 ? sp[0] : sidx == 1 ? sp[1] : sidx == 2 ? sp[2] : sp[3];
            } else if (tid == 4) {
                c = 2 * WS_SHIFT / 4;
#pragma unroll
                for (int q = 0; q < WS_GROUPS; q++) spc += s_c64[q];
            }
            if (c >= c_first && c <= c_last) {
                const float sa = energy(4 * c);
                const float den2 = sa * sb;   // the filter only needs an approximation: rsqrt.approx (2^-22)
                float a = 0.0f, e = 0.0f;
                if (den2 > 1.02f) {
                    a = spc * rsqrt_approx(den2);
                    e = WS_EPS;
                } else if (den2 >= 0.98f) {   // cannot tell which side of 1 the reference's denominator falls
                    a = 0.0f;
                    e = 4.0f;
                }
                ca[c - c_first] = a;
                ce[c - c_first] = e;
                xoff[c - c_first] = 4 * c;
                coff[c - c_first] = -WS_SHIFT + 4 * c;
            }
        }
        __syncthreads();
        const int w1 = ws_decide(D, c_last - c_first + 1, xin, tgt);
        const int best_off = coff[w1];
        const float best_a = ca[w1], best_e = ce[w1];
        __syncthreads();

        // ---- FILTER, fine: best_off-3 .. best_off+3 without best_off, ascending, in bounds (ctts.c:3466-3483)
        int lo = best_off - 3, hi = best_off + 3;
        if (lo < -WS_SHIFT) lo = -WS_SHIFT;
        if (hi > WS_SHIFT) hi = WS_SHIFT;
        // candidate f = thread / 16, split = thread % 16 (24 terms each)
        int nfine = 0;
        {
            const int f = tid >> 4, fs = tid & 15;
            const int off = lo + f;
            float spf = 0.0f;
            const bool live = f <= hi - lo && off != best_off;
            if (live) {
                // interleaved split (term 16 q + fs): the 16 lanes of a candidate read consecutive words
                const float* y = xin + (off + WS_SHIFT) + fs;
                const float* x = tgt + fs;
#pragma unroll
                for (int q = 0; q < WS_OVERLAP / WS_FINE_SPLITS; q++)
                    spf = __fmaf_rn(x[WS_FINE_SPLITS * q], y[WS_FINE_SPLITS * q], spf);
            }
#pragma unroll
            for (int o = 1; o < WS_FINE_SPLITS; o <<= 1) spf += __shfl_xor_sync(0xffffffffu, spf, o);
            if (tid == 0) {   // scan order: the coarse winner first (it is `best`), then the fine candidates ascending
                ca[0] = best_a;
                ce[0] = best_e;
                xoff[0] = best_off + WS_SHIFT;
                coff[0] = best_off;
            }
            __syncthreads();
            // fine candidates in bounds: offsets [flo, fhi] without best_off, in ascending order
            const int flo = lo > -nominal ? lo : -nominal;
            const int fhi = (long long)hi < room ? hi : (int)room;
            if (live && fs == 0 && off >= flo && off <= fhi) {
                const int slot = 1 + (off - flo) - (best_off >= flo && best_off < off ? 1 : 0);
                const int xo = off + WS_SHIFT;
                const float sa = energy(xo);
                const float den2 = sa * sb;
                float a = 0.0f, e = 0.0f;
                if (den2 > 1.02f) {
                    a = spf * rsqrt_approx(den2);
                    e = WS_EPS;
                } else if (den2 >= 0.98f) {
                    e = 4.0f;
                }
                ca[slot] = a;
                ce[slot] = e;
                xoff[slot] = xo;
                coff[slot] = off;
            }
            nfine = fhi >= flo ? fhi - flo + 1 - (best_off >= flo && best_off <= fhi ? 1 : 0) : 0;
        }
        __syncthreads();
        const int w2 = ws_decide(D, 1 + nfine, xin, tgt);
        const int off = coff[w2];
        uint32_t pos = (uint32_t)(nominal + off);
        if (pos + WS_FRAME > n) pos = n - WS_FRAME;
        if (tid == 0) fpos[k] = pos;
        prev_pos = pos;
        __syncthreads();
    }
}

// ---------------------------------------------------------------- speculation: scan + verify

constexpr int WV_THREADS = 256;
constexpr int WV_FRAMES = 24;                                  // frames per tile: (24 + 2) block positions * 19 items = 494 = 2 per thread
constexpr int WV_SPAN = WS_HOP * WV_FRAMES + (WS_RANGE - WS_HOP);   // samples a tile sees: 3840
constexpr int WV_QUADS = WV_SPAN / 4;
constexpr int WV_NB = 3;                                       // tier-1 blocks ...
constexpr int WV_BL = 12;                                      // ... of this many terms (multiple of 4)
constexpr int WV_B0 = 48;                                      // first term of block 0 (multiple of 16: static offsets in the padded layout)
constexpr int WV_BSTEP = WS_HOP;                               // block spacing = the analysis hop: block b of frame k IS block b-1 of frame k+1
constexpr int WV_NJ = WV_FRAMES + WV_NB - 1;                   // block positions a tile's frames use (one hypothesis)
constexpr int WV_DC = 72;                                      // dot products kept per block position: 65 coarse lags, 3 + 3 fine, 1 pad
constexpr int WV_ITEMS = 19;   // per frame: 16 coarse groups of 4, the 65th candidate, fine below, fine above
constexpr int WV_LIST = 768;                                   // tier-2 list entries per tile
constexpr float WV_THR = 2.5e-4f;   // tier 1 rejects when the partial sum of (x^ - t^)^2 exceeds this (5 x the bound, see below)
static_assert(WV_SPAN % 4 == 0 && WV_BL % 4 == 0 && WV_B0 % 16 == 0 && WV_BSTEP % 16 == 0, "float4 alignment, static padded offsets");
static_assert(WV_NJ * 16 + 96 <= 2 * WV_THREADS && WV_NJ <= 32 && (WV_NJ * 16) % 32 == 0, "two items per thread, kinds on warp boundaries");
static_assert(WV_B0 + (WV_NB - 1) * WV_BSTEP + WV_BL <= WS_OVERLAP, "blocks inside the window");

// Shared-memory layout of the staged samples: 4 floats of padding after every 16, so that the 16-float
// segments consecutive lanes start their float4 loads at (one group of 4 coarse candidates per lane) fall
// on different banks (lane stride 80 bytes: 8 lanes of a quarter warp cover all 32 banks once).
__host__ __device__ constexpr int wv_phys(int i) { return i + ((i >> 4) << 2); }

struct WvSmem {
    alignas(16) float xs[wv_phys(WV_SPAN + 16)];   // samples of the tile as floats (zero past the end of the input), padded layout
    alignas(16) unsigned long long PQ[WV_QUADS + 1];   // PQ[j] = sum_{i < 4j} x[i]^2, exact
    alignas(16) float e4[WV_QUADS + 4];        // energy of quad j (float of the exact integer)
    alignas(16) float W4[WV_QUADS];            // window energy at 4j (float of the exact integer)
    alignas(16) float G4[WV_QUADS];            // energy of the tier-1 blocks of the window at 4j
    alignas(16) float D[WV_NJ * WV_DC];        // partial dot products per (block position, lag)
    float Sl[WV_NJ * 8];                       // per block position: block energy of the fine lags minus that of the lag they are slid from
    float Et[WV_FRAMES], Bt[WV_FRAMES];        // target: energy, block energy / energy
    uint32_t has_hyp[2];                       // tile has live frames under hypothesis -128 / 0
    int hyp[WV_FRAMES];                        // speculated previous offset; INT_MIN: frame needs no check
    int room[WV_FRAMES];                       // largest offset in bounds
    unsigned long long wtot[WV_THREADS / 32];
    uint32_t list[WV_LIST];                    // (frame << 16) | (offset + 128)
    uint32_t n_list;
    uint32_t bad;                              // smallest bad frame of the tile
};

// per stretched utterance (one warp each): frame count, the first frame whose target under offset 0
// is digital silence, and the reset of the per-run state
__global__ void __launch_bounds__(256) wsola_scan_kernel(const WsolaArgs A) {
    const uint32_t w = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (w >= A.task_count) return;
    const uint32_t ti = A.task_first + w;
    const StretchTask task = A.tasks[ti];
    const uint32_t n = A.pre_counts[task.utt];
    const int16_t* in = A.pre + task.pre_off;
    uint32_t frames = n >= WS_FRAME ? (n - WS_FRAME) / WS_HOP + 1 : 0;
    if (frames > task.max_frames) frames = task.max_frames;
    uint32_t F = 0xffffffffu;
    if (A.speculate) {
        // target of frame k when frame k-1 sits at its nominal position: in[128 k .. 128 k + 384)
        for (uint32_t k0 = 1; k0 < frames && F == 0xffffffffu; k0 += 32) {
            const uint32_t k = k0 + (uint32_t)lane;
            bool zero = k < frames;
            if (zero) {
                const int4* v = reinterpret_cast<const int4*>(in + (size_t)k * WS_HOP);   // pre_off and 128 k are multiples of 8
                for (int i = 0; i < WS_OVERLAP / 8 && zero; i++) {
                    const int4 q = v[i];
                    zero = (q.x | q.y | q.z | q.w) == 0;
                }
            }
            const uint32_t m = __ballot_sync(0xffffffffu, zero);
            if (m) F = k0 + (uint32_t)__ffs(m) - 1u;
        }
    }
    if (lane == 0) {
        A.n_frames[ti] = frames;
        A.n_frames[A.n_tasks + ti] = 0;
        A.tier2[ti] = 0;
        A.first_silent[ti] = F;
        A.first_bad[ti] = A.speculate ? 0xffffffffu : 1u;
        if (frames) A.frame_pos[task.pos_off] = 0;
    }
}

// tier 1, one item: the WV_BL cross terms of ONE block for 4 candidates (candidate blocks start at
// yb + SHIFT + STRIDE * j, yb a multiple of 16 in the tile) against the target block at tb (a multiple of 16):
// one FMA per term, operand loads shared by the 4 candidates.  slide[j] (STRIDE == 1): energy of candidate
// block j minus that of candidate block 0.
template <int STRIDE, int SHIFT>
__device__ __forceinline__ void wv_block_dots(const float* __restrict__ xs, int tb, int yb, float (&d)[4], float (&slide)[4]) {
    d[0] = d[1] = d[2] = d[3] = 0.0f;
    slide[0] = slide[1] = slide[2] = slide[3] = 0.0f;
    constexpr int TC = WV_BL / 4;                                 // target chunks
    constexpr int YC = STRIDE == 4 ? TC + 3 : TC + 1;             // candidate chunks
    const float* tp = xs + wv_phys(tb);
    const float* yp = xs + wv_phys(yb);
    float y[4 * YC];
#pragma unroll
    for (int m = 0; m < YC; m++) {
        const float4 v = *reinterpret_cast<const float4*>(yp + wv_phys(SHIFT + 4 * m));
        y[4 * m] = v.x; y[4 * m + 1] = v.y; y[4 * m + 2] = v.z; y[4 * m + 3] = v.w;
    }
#pragma unroll
    for (int m = 0; m < TC; m++) {
        const float4 t = *reinterpret_cast<const float4*>(tp + wv_phys(4 * m));
#pragma unroll
        for (int j = 0; j < 4; j++) {
            d[j] = __fmaf_rn(t.x, y[STRIDE * j + 4 * m], d[j]);
            d[j] = __fmaf_rn(t.y, y[STRIDE * j + 4 * m + 1], d[j]);
            d[j] = __fmaf_rn(t.z, y[STRIDE * j + 4 * m + 2], d[j]);
            d[j] = __fmaf_rn(t.w, y[STRIDE * j + 4 * m + 3], d[j]);
        }
    }
    if (STRIDE == 1) {
        float acc = 0.0f;
#pragma unroll
        for (int j = 1; j < 4; j++) {
            acc += y[WV_BL + j - 1] * y[WV_BL + j - 1] - y[j - 1] * y[j - 1];
            slide[j] = acc;
        }
    }
}

// Verification of the speculated offsets, one tile of WV_FRAMES frames of one utterance per CTA
// (blockIdx.y = task of the launch, blockIdx.x = tile).
//
// Frame k, speculated previous offset h (0 up to and including frame F, -128 after it): the target is
// in[128 k + h .. + 384).  If it is all zero the frame's offset is -128 (see the header).  Otherwise the
// candidate at offset h is the target itself and scores exactly 1.0f, and the frame's offset is h
// provided every other candidate the reference would score -- coarse offsets -128, -124 .. 128 and
// fine offsets h-3 .. h+3, in bounds -- scores < 1.0f.
//
// Error budget of tier 1.  With x^ = x/|x|, t^ = t/|t| (exact real arithmetic) the true correlation is
// rho = 1 - |x^ - t^|^2 / 2 <= 1 - S/2, S = sum over the blocks of (x^_i - t^_i)^2 = A + B - 2C,
// A = E_blocks(x)/E(x), B = E_blocks(t)/E(t) (window energies exact integers from the prefix sum, block
// energies float sums of <= 36 exact squares, approximate reciprocal: relative error < 3e-6 each incl. the
// slid energies of the fine candidates), C = dot/sqrt(E(x)E(t)) (36 FMA terms in three partial sums, rsqrt.approx:
// |error| <= (gamma_14 + 2^-22 + 5u) sqrt(AB) < 1.5e-6).  So |S~ - S| < 1e-5 (A, B <= 1), and the
// reference's score r satisfies |r - rho| <= 2e-5, hence r <= 1 - S~/2 + 5e-6 + 2e-5 < 1
// whenever S~ > 5e-5.  WV_THR is 2.5e-4.
__global__ void __launch_bounds__(WV_THREADS, 4) wsola_verify_kernel(const WsolaArgs A) {
    extern __shared__ __align__(16) unsigned char wv_raw[];
    WvSmem& sm = *reinterpret_cast<WvSmem*>(wv_raw);
    const uint32_t ti = A.task_first + blockIdx.y;
    const uint32_t frames = A.n_frames[ti];
    const uint32_t k0 = 1u + blockIdx.x * WV_FRAMES;        // first frame of the tile
    if (k0 >= frames) return;
    const StretchTask task = A.tasks[ti];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t n = A.pre_counts[task.utt];
    const int16_t* in = A.pre + task.pre_off;
    const uint32_t F = A.first_silent[ti];
    const uint32_t base = (k0 - 1u) * WS_HOP;               // tile sample 0 = input sample base

    // ---- stage the tile as floats; exact prefix sum of squares at every 4th sample.  Thread t owns the 16
    //      samples 16 t .. 16 t + 15 (two 16-byte loads; its padded float4 stores are conflict free).
    if (tid == 0) {
        sm.n_list = 0;
        sm.bad = 0xffffffffu;
        sm.PQ[0] = 0ull;
        sm.has_hyp[0] = sm.has_hyp[1] = 0u;
    }
    if (tid >= WV_QUADS / 4 && tid < WV_QUADS / 4 + 4) {
        const int z = tid - WV_QUADS / 4;
        *reinterpret_cast<float4*>(sm.xs + wv_phys(WV_SPAN + 4 * z)) = make_float4(0.f, 0.f, 0.f, 0.f);
        sm.e4[WV_QUADS + z] = 0.0f;
    }
    static_assert(WV_QUADS % 4 == 0 && WV_QUADS / 4 + 4 <= WV_THREADS, "one thread per 16 samples");
    {
        unsigned long long qs[4] = {0ull, 0ull, 0ull, 0ull};   // inclusive sums of the thread's quads
        if (tid < WV_QUADS / 4) {
            const uint32_t p = base + 16u * (uint32_t)tid;
            int4 lo = make_int4(0, 0, 0, 0), hi = make_int4(0, 0, 0, 0);
            if (p + 16u <= n) {
                lo = *reinterpret_cast<const int4*>(in + p);
                hi = *reinterpret_cast<const int4*>(in + p + 8);
            } else if (p < n) {
                int16_t* e = reinterpret_cast<int16_t*>(&lo);
                int16_t* g = reinterpret_cast<int16_t*>(&hi);
                for (uint32_t i = 0; i < 8; i++) {
                    if (p + i < n) e[i] = in[p + i];
                    if (p + 8 + i < n) g[i] = in[p + 8 + i];
                }
            }
            const uint32_t w[8] = {(uint32_t)lo.x, (uint32_t)lo.y, (uint32_t)lo.z, (uint32_t)lo.w,
                                   (uint32_t)hi.x, (uint32_t)hi.y, (uint32_t)hi.z, (uint32_t)hi.w};
            float4 e4v;
            float* e4p = reinterpret_cast<float*>(&e4v);
            unsigned long long run = 0ull;
#pragma unroll
            for (int m = 0; m < 4; m++) {
                const int v0 = (int)(short)(w[2 * m] & 0xffffu), v1 = (int)(short)(w[2 * m] >> 16);
                const int v2 = (int)(short)(w[2 * m + 1] & 0xffffu), v3 = (int)(short)(w[2 * m + 1] >> 16);
                *reinterpret_cast<float4*>(sm.xs + wv_phys(16 * tid + 4 * m)) = make_float4((float)v0, (float)v1, (float)v2, (float)v3);
                const unsigned long long s3 = (unsigned long long)(uint32_t)(v0 * v0) + (uint32_t)(v1 * v1) +
                                              (unsigned long long)(uint32_t)(v2 * v2) + (uint32_t)(v3 * v3);
                e4p[m] = ws_u64_to_float(s3);   // < 2^32: two roundings of 2^-24
                run += s3;
                qs[m] = run;
            }
            *reinterpret_cast<float4*>(sm.e4 + 4 * tid) = e4v;
        }
        unsigned long long inc = qs[3];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) sm.wtot[warp] = inc;
        __syncthreads();
        unsigned long long pre = inc - qs[3];   // exclusive prefix at the thread's first sample
#pragma unroll
        for (int w = 0; w < WV_THREADS / 32; w++)
            if (w < warp) pre += sm.wtot[w];
        if (tid < WV_QUADS / 4) {
            ulonglong2 a, b2;
            a.x = pre + qs[0]; a.y = pre + qs[1];
            b2.x = pre + qs[2]; b2.y = pre + qs[3];
            sm.PQ[4 * tid + 1] = a.x;   // (PQ + 4 t + 1 is 8-byte aligned only)
            sm.PQ[4 * tid + 2] = a.y;
            sm.PQ[4 * tid + 3] = b2.x;
            sm.PQ[4 * tid + 4] = b2.y;
        }
        __syncthreads();
    }

    // ---- window energies (exact) and tier-1 block energies at every 4th sample
    for (int j = tid; j < WV_QUADS; j += WV_THREADS) {
        float w = 0.0f, g = 0.0f;
        if (4 * j + WS_OVERLAP <= WV_SPAN) {
            w = ws_u64_to_float(sm.PQ[j + WS_OVERLAP / 4] - sm.PQ[j]);
#pragma unroll
            for (int b = 0; b < WV_NB; b++)
#pragma unroll
                for (int m = 0; m < WV_BL / 4; m++) g += sm.e4[j + (WV_B0 + b * WV_BSTEP) / 4 + m];
        }
        sm.W4[j] = w;
        sm.G4[j] = g;
    }
    __syncthreads();
    if (tid < WV_FRAMES) {
        const uint32_t k = k0 + (uint32_t)tid;
        int h = INT_MIN;
        if (k < frames) {
            h = k <= F ? 0 : -WS_SHIFT;
            const int ts = WS_HOP * tid + WS_SHIFT + h;
            const float et = sm.W4[ts >> 2];            // float of the exact integer: 0 iff all zero
            uint32_t pos = k * WS_HOP + (uint32_t)h;
            if (et == 0.0f) {   // digital silence behind the frame: offset -128, nothing to verify
                pos = k * WS_HOP - WS_SHIFT;
                h = INT_MIN;
            } else {
                sm.Et[tid] = et;
                sm.Bt[tid] = __fdividef(sm.G4[ts >> 2], et);
                sm.has_hyp[h == 0 ? 1 : 0] = 1u;
            }
            if (A.force_bad && k % A.force_bad == 0) atomicMin(&sm.bad, k);
            A.frame_pos[task.pos_off + k] = pos;
            const long long room = (long long)n - WS_FRAME - (long long)k * WS_HOP;   // >= 0
            sm.room[tid] = room > WS_SHIFT ? WS_SHIFT : (int)room;
        }
        sm.hyp[tid] = h;
    }
    __syncthreads();

    // ---- tier 1.  The blocks are one analysis hop apart, so the partial dot product of block b of frame k
    //      at a given lag IS that of block b-1 of frame k+1 (same samples, as long as both frames have the
    //      same speculated offset): they are formed once per (block position, lag) -- D -- and every
    //      (frame, candidate) then adds its three.  A tile that holds the one frame where the speculated
    //      offset changes does this twice, once per offset.
    for (int pass = 0; pass < 2; pass++) {
        if (!sm.has_hyp[pass]) continue;                          // CTA-uniform
        const int hp = pass ? 0 : -WS_SHIFT;                      // the speculated offset of this pass
        const int j_lo = pass;                                    // first block position used: frame f, block b sits at f + b + j_lo
        // items are ordered by KIND so that a warp runs one kind of code: 16 coarse groups per block position
        // (WV_NJ * 16 = 13 warps), then -- 32 slots each -- the 65th candidate, the fine lags below, the fine lags above
        for (int it = tid; it < WV_NJ * 16 + 96; it += WV_THREADS) {
            float d[4], slide[4];
            if (it < WV_NJ * 16) {                                // lags 16 g - 128 - hp + {0, 4, 8, 12}
                const int jl = it >> 4, g = it & 15;
                const int tb = WS_HOP * (jl + j_lo) + WV_B0;      // target block (multiple of 16)
                wv_block_dots<4, 0>(sm.xs, tb, tb + 16 * g - WS_SHIFT - hp, d, slide);
                *reinterpret_cast<float4*>(sm.D + jl * WV_DC + 4 * g) = make_float4(d[0], d[1], d[2], d[3]);
                continue;
            }
            const int kind = (it - WV_NJ * 16) >> 5, jl = (it - WV_NJ * 16) & 31;
            if (jl >= WV_NJ) continue;
            const int tb = WS_HOP * (jl + j_lo) + WV_B0;
            float* Dj = sm.D + jl * WV_DC;
            if (kind == 0) {                                      // the 65th coarse candidate alone
                wv_block_dots<4, 0>(sm.xs, tb, tb + WS_SHIFT - hp, d, slide);
                Dj[64] = d[0];
            } else if (kind == 1) {                               // fine, below: lags -3, -2, -1 (slid from -4)
                wv_block_dots<1, 12>(sm.xs, tb, tb - 16, d, slide);
                Dj[65] = d[1]; Dj[66] = d[2]; Dj[67] = d[3];
                sm.Sl[8 * jl] = slide[1]; sm.Sl[8 * jl + 1] = slide[2]; sm.Sl[8 * jl + 2] = slide[3];
            } else {                                              // fine, above: lags 1, 2, 3 (slid from 0)
                wv_block_dots<1, 0>(sm.xs, tb, tb, d, slide);
                Dj[68] = d[1]; Dj[69] = d[2]; Dj[70] = d[3];
                sm.Sl[8 * jl + 3] = slide[1]; sm.Sl[8 * jl + 4] = slide[2]; sm.Sl[8 * jl + 5] = slide[3];
            }
        }
        __syncthreads();
        auto inconclusive = [&](int f, int o) {
            const uint32_t slot = atomicAdd(&sm.n_list, 1u);
            if (slot < (uint32_t)WV_LIST) sm.list[slot] = ((uint32_t)f << 16) | (uint32_t)(o + WS_SHIFT);
            else atomicMin(&sm.bad, k0 + (uint32_t)f);             // no room to look closer: let the chain walk decide
        };
        // (frame, group of 4 coarse candidates): three float4 of partial dots, one each of window / block energies
        for (int it = tid; it < WV_FRAMES * 16; it += WV_THREADS) {
            const int f = it >> 4, g = it & 15;
            if (sm.hyp[f] != hp) continue;
            const int room = sm.room[f];
            const float et = sm.Et[f], bt = sm.Bt[f];
            const float4 d0 = *reinterpret_cast<const float4*>(sm.D + f * WV_DC + 4 * g);
            const float4 d1 = *reinterpret_cast<const float4*>(sm.D + (f + 1) * WV_DC + 4 * g);
            const float4 d2 = *reinterpret_cast<const float4*>(sm.D + (f + 2) * WV_DC + 4 * g);
            const int q = (WS_HOP * f + 16 * g) >> 2;             // window of offset -128 + 16 g: tile sample 128 f + 16 g
            const float4 w4 = *reinterpret_cast<const float4*>(sm.W4 + q);
            const float4 g4 = *reinterpret_cast<const float4*>(sm.G4 + q);
            const float dot[4] = {d0.x + d1.x + d2.x, d0.y + d1.y + d2.y, d0.z + d1.z + d2.z, d0.w + d1.w + d2.w};
            const float w[4] = {w4.x, w4.y, w4.z, w4.w}, gb[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int o = -WS_SHIFT + 16 * g + 4 * j;
                if (o == hp || o > room || w[j] == 0.0f) continue;   // the target itself / out of bounds / score exactly 0 (ctts.c:3426)
                const float sres = __fdividef(gb[j], w[j]) + bt - 2.0f * dot[j] * rsqrt_approx(w[j] * et);
                if (!(sres > WV_THR)) inconclusive(f, o);
            }
        }
        // (frame, one of: the 65th coarse candidate, 3 fine lags below, 3 above): 8 slots per frame
        for (int it = tid; it < WV_FRAMES * 8; it += WV_THREADS) {
            const int f = it >> 3, c7 = it & 7;
            if (sm.hyp[f] != hp || c7 == 7) continue;
            const int ts = WS_HOP * f + WS_SHIFT + hp;            // the target in the tile
            const int room = sm.room[f];
            const int c = 64 + c7;
            int o;
            float w, gb;
            if (c7 == 0) {
                o = WS_SHIFT;
                if (o > room) continue;
                const int q = (WS_HOP * f + 2 * WS_SHIFT) >> 2;
                w = sm.W4[q];
                gb = sm.G4[q];
                if (w == 0.0f) continue;
            } else {
                const int below = c7 < 4, j = below ? c7 : c7 - 3;               // slid by j = 1..3 samples
                o = hp + (below ? j - 4 : j);
                if (o < -WS_SHIFT || o > WS_SHIFT || o > room) continue;
                const int p0 = ts + (below ? -4 : 0);             // the window the energies are slid from (multiple of 4)
                const float wb = sm.W4[p0 >> 2];
                const float4 x0 = *reinterpret_cast<const float4*>(sm.xs + wv_phys(p0));
                const float4 x1 = *reinterpret_cast<const float4*>(sm.xs + wv_phys(p0 + WS_OVERLAP));
                w = wb + (x1.x * x1.x - x0.x * x0.x);
                if (j > 1) w += x1.y * x1.y - x0.y * x0.y;
                if (j > 2) w += x1.z * x1.z - x0.z * x0.z;
                const int k = (below ? 0 : 3) + j - 1;
                gb = sm.G4[p0 >> 2] + sm.Sl[8 * f + k] + sm.Sl[8 * (f + 1) + k] + sm.Sl[8 * (f + 2) + k];
                // a slid window energy is trusted only while it has not lost most of the energy it was slid from
                // (cancellation: its relative error is then < 8 * 3 * 2^-24); else the candidate is looked at closely
                if (!(w >= 1.0f && w >= 0.125f * wb)) w = -1.0f;
            }
            float sres = -1.0f;
            if (w > 0.0f) {
                const float dot = sm.D[f * WV_DC + c] + sm.D[(f + 1) * WV_DC + c] + sm.D[(f + 2) * WV_DC + c];
                sres = __fdividef(gb, w) + sm.Bt[f] - 2.0f * dot * rsqrt_approx(w * sm.Et[f]);
            }
            if (!(sres > WV_THR)) inconclusive(f, o);
        }
        __syncthreads();
    }

    // ---- tiers 2 and 3: one warp per candidate tier 1 could not reject
    const uint32_t n_list = min(sm.n_list, (uint32_t)WV_LIST);
    uint32_t n_exact = 0;
    for (uint32_t e = warp; e < n_list; e += WV_THREADS / 32) {
        const uint32_t ent = sm.list[e];
        const int f = (int)(ent >> 16), xo = (int)(ent & 0xffffu);   // xo = offset + 128: window start inside the view
        const int tg = WS_HOP * f + WS_SHIFT + sm.hyp[f], xc = WS_HOP * f + xo;
        float dot = 0.0f, ea = 0.0f;
        if ((xc & 3) == 0) {   // a coarse candidate: 16-byte vectors (the padding never splits an aligned quad)
#pragma unroll
            for (int q = 0; q < WS_OVERLAP / 128; q++) {
                const float4 a = *reinterpret_cast<const float4*>(sm.xs + wv_phys(xc + 4 * lane + 128 * q));
                const float4 t = *reinterpret_cast<const float4*>(sm.xs + wv_phys(tg + 4 * lane + 128 * q));
                dot = __fmaf_rn(t.x, a.x, dot); dot = __fmaf_rn(t.y, a.y, dot);
                dot = __fmaf_rn(t.z, a.z, dot); dot = __fmaf_rn(t.w, a.w, dot);
            }
        } else {
#pragma unroll
            for (int q = 0; q < WS_OVERLAP / 32; q++) {
                const float a = sm.xs[wv_phys(xc + lane + 32 * q)];
                dot = __fmaf_rn(sm.xs[wv_phys(tg + lane + 32 * q)], a, dot);
                ea = __fmaf_rn(a, a, ea);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            dot += __shfl_xor_sync(0xffffffffu, dot, o);
            ea += __shfl_xor_sync(0xffffffffu, ea, o);
        }
        // the candidate's energy: exact where the window starts on a quad, else a float sum of exact squares
        // (relative error < 2e-6, far inside the filter's eps)
        const float sa = (xc & 3) == 0 ? sm.W4[xc >> 2] : ea;
        if (sa == 0.0f) continue;                        // all zero: the score is exactly 0
        // the full FMA filter score: within 2e-5 of the reference's (eps = 1e-4 as in the chain kernel)
        const float a = dot * rsqrt_approx(sa * sm.Et[f]);
        if (a + WS_EPS < 1.0f) continue;
        // tier 3: the reference's loop; lanes 0..2 form sum_prod, sum_sq1, sum_sq2 (ctts.c:3411-3413)
        float acc = 0.0f;
        if (lane < 3) {
            const int p = lane == 2 ? tg : xc, q = lane == 1 ? xc : tg;
#pragma unroll 2
            for (int m = 0; m < WS_OVERLAP / 4; m++) {
                const float p0 = sm.xs[wv_phys(p + 4 * m)], p1 = sm.xs[wv_phys(p + 4 * m + 1)], p2 = sm.xs[wv_phys(p + 4 * m + 2)],
                            p3 = sm.xs[wv_phys(p + 4 * m + 3)];
                const float q0 = sm.xs[wv_phys(q + 4 * m)], q1 = sm.xs[wv_phys(q + 4 * m + 1)], q2 = sm.xs[wv_phys(q + 4 * m + 2)],
                            q3 = sm.xs[wv_phys(q + 4 * m + 3)];
                acc += p0 * q0 + p1 * q1 + p2 * q2 + p3 * q3;
            }
        }
        const float sp = __shfl_sync(0xffffffffu, acc, 0), s1 = __shfl_sync(0xffffffffu, acc, 1), s2 = __shfl_sync(0xffffffffu, acc, 2);
        const float den = sqrtf(s1 * s2);
        const float r = den < 1.0f ? 0.0f : sp / den;
        n_exact++;
        // a candidate that reaches 1.0f ties with (or beats) the speculated one: scan order decides,
        // which is the chain walk's business
        if (!(r < 1.0f) && lane == 0) atomicMin(&sm.bad, k0 + (uint32_t)f);
    }
    if (lane == 0 && n_exact) atomicAdd(A.n_frames + A.n_tasks + ti, n_exact);
    __syncthreads();
    if (tid == 0) {
        if (n_list) atomicAdd(A.tier2 + ti, n_list);
        if (sm.bad != 0xffffffffu) atomicMin(A.first_bad + ti, sm.bad);
    }
}

// Output samples one OLA block covers for synthesis hop `hop` (host and device must agree).
// 64 <= hop <= OLA_THREADS (every hop the C-ABI can produce: hop = (size_t)(128 / speed), speed in
// [0.5, 2]): OLA_THREADS / hop row groups of OLA_SPT rows of hop samples; otherwise plain tiles.
__host__ __device__ inline uint32_t ola_block_span(uint32_t hop) {
    return hop >= 64 && hop <= (uint32_t)OLA_THREADS ? ((uint32_t)OLA_THREADS / hop) * (uint32_t)OLA_SPT * hop
                                                     : (uint32_t)OLA_THREADS * (uint32_t)OLA_SPT;
}

// Overlap-add in gather form (ctts.c:3563-3606).  Output sample j = row * hop + c (row = j / hop)
// receives frame k = row - q at window index i = q * hop + c, for every q with i < 512 and
// 0 <= k < frames.  A thread owns one residue c and OLA_SPT consecutive rows: the window value
// w[q * hop + c] is the same for all of its outputs, and walking q downwards adds the frames of
// every output in ascending frame order, which is the order the reference accumulates the float
// norm in (the int16 accumulator wraps, ctts.c:3577: it is the 32-bit sum mod 2^16).
constexpr int OLA_STAGE = WS_HOP * (OLA_THREADS / 64) * OLA_SPT + 1536;   // input samples an interior block can touch (hop >= 64)

__global__ void __launch_bounds__(OLA_THREADS, 4) wsola_ola_kernel(const WsolaArgs A) {
    __shared__ int s_last[OLA_THREADS / 32];
    __shared__ float win[WS_FRAME];
    __shared__ uint32_t sfp[(OLA_THREADS / 64) * OLA_SPT + 8];
    __shared__ __align__(16) float xin[OLA_STAGE + 8];
    const int tid = threadIdx.x;
    for (int i = tid; i < WS_FRAME; i += OLA_THREADS) win[i] = __ldg(A.hann512 + i);
    const uint32_t ti = A.ola_block_task[blockIdx.x];
    const StretchTask task = A.tasks[ti];
    const uint32_t frames = A.n_frames[ti];
    if (frames == 0) return;   // CTA-uniform
    const uint32_t hop = task.hop;
    const uint32_t used = (frames - 1) * hop + WS_FRAME;
    const uint32_t lim = used < task.out_cap ? used : task.out_cap;
    const int16_t* in = A.pre + task.pre_off;
    const uint32_t* fpos = A.frame_pos + task.pos_off;
    int16_t* out = A.out + task.out_off;
    const uint32_t first = A.ola_block_first[blockIdx.x];
    int last_nz = -1;
    const uint32_t groups_ = hop >= 64 && hop <= (uint32_t)OLA_THREADS ? (uint32_t)OLA_THREADS / hop : 0u;
    const uint32_t row0_ = groups_ ? first / hop : 0u;
    // INTERIOR block: every output is covered by whole frames that all exist (no start / end effects), so
    // the norm of an output depends on its residue alone and the frames' input span is bounded: the span is
    // staged once as floats (an input sample feeds 4 frames x up to 2 outputs), the per-thread norm and its
    // reciprocal are loop invariants, and a contribution costs a shared load, a multiply, a truncation, an add.
    if (groups_ && row0_ >= 7 && row0_ + groups_ * OLA_SPT - 1 <= frames - 1 && first + groups_ * OLA_SPT * hop <= lim) {
        const uint32_t groups = groups_, row0 = row0_;
        const uint32_t n_in = A.pre_counts[task.utt];
        for (uint32_t i = tid; i < groups * OLA_SPT + 7; i += OLA_THREADS) sfp[i] = __ldg(fpos + (row0 - 7 + i));
        // frame k sits at 128 k + offset, |offset| <= 128 (and never past n - 512)
        const uint32_t s0 = (row0 - 7) * WS_HOP - WS_SHIFT;   // row0 >= 7, multiple of 8
        const uint32_t span = WS_HOP * (groups * OLA_SPT + 6) + 2 * WS_SHIFT + WS_FRAME;   // <= OLA_STAGE
        for (uint32_t v = tid; v < span / 8; v += OLA_THREADS) {
            const uint32_t p = s0 + 8 * v;
            int4 q = make_int4(0, 0, 0, 0);
            if (p + 8 <= n_in) q = *reinterpret_cast<const int4*>(in + p);
            else if (p < n_in) {
                int16_t* e = reinterpret_cast<int16_t*>(&q);
                for (uint32_t k = 0; k < 8 && p + k < n_in; k++) e[k] = in[p + k];
            }
            float4 a, b;
            a.x = (float)(short)((uint32_t)q.x & 0xffffu); a.y = (float)(short)((uint32_t)q.x >> 16);
            a.z = (float)(short)((uint32_t)q.y & 0xffffu); a.w = (float)(short)((uint32_t)q.y >> 16);
            b.x = (float)(short)((uint32_t)q.z & 0xffffu); b.y = (float)(short)((uint32_t)q.z >> 16);
            b.z = (float)(short)((uint32_t)q.w & 0xffffu); b.w = (float)(short)((uint32_t)q.w >> 16);
            *(reinterpret_cast<float4*>(xin) + 2 * v) = a;
            *(reinterpret_cast<float4*>(xin) + 2 * v + 1) = b;
        }
        __syncthreads();
        const uint32_t g = (uint32_t)tid / hop, c = (uint32_t)tid - g * hop;
        if (g < groups) {
            const uint32_t rowbase = row0 + g * OLA_SPT;
            // positions (relative to the staged span) of frames rowbase - 7 .. rowbase + OLA_SPT - 1
            int fpr[OLA_SPT + OLA_QMAX - 1];
            bool inside = true;
#pragma unroll
            for (int k = 0; k < OLA_SPT + OLA_QMAX - 1; k++) {
                fpr[k] = (int)(sfp[g * OLA_SPT + k] - s0);
                inside &= fpr[k] >= 0 && fpr[k] + WS_FRAME <= (int)span;
            }
            const int qmax = (int)((WS_FRAME - 1) / hop);   // <= 7
            float nrm = 0.0f;
            int acc[OLA_SPT];
#pragma unroll
            for (int r = 0; r < OLA_SPT; r++) acc[r] = 0;
#pragma unroll
            for (int q = OLA_QMAX - 1; q >= 0; q--) {
                const uint32_t i = (uint32_t)q * hop + c;
                if (q <= qmax && i < (uint32_t)WS_FRAME) {
                    const float wv = win[i];
                    nrm += wv;                       // ascending frame order, as the reference accumulates it (ctts.c:3578)
                    if (inside) {
#pragma unroll
                        for (int r = 0; r < OLA_SPT; r++)   // |x * w| <= 32767: the int32 truncation is the int16 value
                            acc[r] += (int)(xin[fpr[r - q + OLA_QMAX - 1] + (int)i] * wv);
                    } else {                          // (a frame outside the analytic span: cannot happen, kept for safety)
#pragma unroll
                        for (int r = 0; r < OLA_SPT; r++)
                            acc[r] += (int)((float)in[sfp[g * OLA_SPT + r - q + OLA_QMAX - 1] + i] * wv);
                    }
                }
            }
            const bool norm_ok = nrm > 0.01f;
            const float rn = recip_for_div(norm_ok ? nrm : 1.0f);
#pragma unroll
            for (int r = 0; r < OLA_SPT; r++) {
                const uint32_t j = (rowbase + (uint32_t)r) * hop + c;
                const int a16 = (int)(int16_t)acc[r];
                const int y = norm_ok ? cvt_sat_s16(div_by((float)a16, nrm, rn)) : a16;
                out[j] = (int16_t)y;
                if (y != 0) last_nz = (int)j;
            }
        }
    } else
    if (hop >= 64 && hop <= (uint32_t)OLA_THREADS) {
        const uint32_t groups = (uint32_t)OLA_THREADS / hop;
        const uint32_t row0 = first / hop;   // exact: first is a multiple of the block span
        // positions of frames row0 - 7 .. row0 + groups * OLA_SPT - 1 (0 where there is no such frame)
        for (uint32_t i = tid; i < groups * OLA_SPT + 7; i += OLA_THREADS) {
            const long long k = (long long)row0 - 7 + i;
            sfp[i] = (k >= 0 && k < (long long)frames) ? __ldg(fpos + k) : 0u;
        }
        __syncthreads();
        const uint32_t g = (uint32_t)tid / hop, c = (uint32_t)tid - g * hop;
        if (g < groups && first < lim) {
            const uint32_t rowbase = row0 + g * OLA_SPT;
            const uint32_t kmax = frames - 1;
            int acc[OLA_SPT];
            float nrm[OLA_SPT];
#pragma unroll
            for (int r = 0; r < OLA_SPT; r++) {
                acc[r] = 0;
                nrm[r] = 0.0f;
            }
            for (int q = (int)((WS_FRAME - 1) / hop); q >= 0; q--) {   // CTA-uniform trip count, <= 8
                const uint32_t i = (uint32_t)q * hop + c;
                if (i < (uint32_t)WS_FRAME) {
                    const float wv = win[i];
                    const uint32_t* fp = sfp + (g * OLA_SPT + 7u - (uint32_t)q);   // frame rowbase - q
#pragma unroll
                    for (int r = 0; r < OLA_SPT; r++) {
                        const uint32_t k = rowbase + (uint32_t)r - (uint32_t)q;    // wraps below frame 0
                        if (k <= kmax) {
                            acc[r] += (int)f2s((float)in[fp[r] + i] * wv);
                            nrm[r] += wv;
                        }
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < OLA_SPT; r++) {
                const uint32_t j = (rowbase + (uint32_t)r) * hop + c;
                if (j < lim) {
                    const int16_t a = (int16_t)acc[r];
                    int16_t y = a;
                    if (nrm[r] > 0.01f) y = f2s(clamp16f((float)a / nrm[r]));
                    out[j] = y;
                    if (y != 0) last_nz = (int)j;
                }
            }
        }
    } else {
        __syncthreads();
        for (int r = 0; r < OLA_SPT; r++) {
            const uint32_t j = first + (uint32_t)r * OLA_THREADS + (uint32_t)tid;
            if (j >= lim) continue;
            uint32_t k_hi = j / hop;
            if (k_hi > frames - 1) k_hi = frames - 1;
            const uint32_t k_lo = j < WS_FRAME ? 0u : (j - WS_FRAME) / hop + 1u;
            int16_t acc = 0;
            float norm = 0.0f;
            for (uint32_t k = k_lo; k <= k_hi; k++) {
                const uint32_t i = j - k * hop;
                const float wv = win[i];
                acc = (int16_t)(acc + f2s((float)in[__ldg(fpos + k) + i] * wv));
                norm += wv;
            }
            int16_t y = acc;
            if (norm > 0.01f) y = f2s(clamp16f((float)acc / norm));
            out[j] = y;
            if (y != 0) last_nz = (int)j;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        int other = __shfl_xor_sync(0xffffffffu, last_nz, o);
        last_nz = other > last_nz ? other : last_nz;
    }
    if (lane_id() == 0) s_last[warp_id()] = last_nz;
    __syncthreads();
    if (tid == 0) {
        int m = -1;
        for (int w = 0; w < OLA_THREADS / 32; w++) m = s_last[w] > m ? s_last[w] : m;
        if (m >= 0) atomicMax(A.out_counts + task.utt, (uint32_t)(m + 1));
    }
}

}  // namespace ctts
