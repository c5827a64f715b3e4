// assemble.cuh -- the audio-assembly kernel.
//
// Decomposition.  The reference is one sequential program per utterance over a
// growing buffer (ctts.c:3689-3904).  What one word region (the samples between
// two word marks) does depends on earlier regions only through
//   (1) the absolute sample count at its start (the `count/2`, `count` clamps of
//       ctts.c:1985-1987, :1736, :3319 and the `count == 0` tests), and
//   (2) rarely, the last few thousand finished samples (an analysis / crossfade /
//       fade window that reaches back past the word start).
// So the parallel unit here is the REGION TASK: one CTA assembles one region (or
// a run of tiny ones) entirely in shared memory, and only at the end -- or at
// the first op whose decision really needs (1) or (2) -- waits for its
// predecessor's published inclusive sample count (a decoupled look-back chain,
// one 64-bit word per task).  The finished region is then streamed to its final
// position in the utterance's HBM slot with 16-byte stores.  Tasks are handed
// out through an atomic ticket in region-major order (region r of every
// utterance before region r+1 of any), so predecessors are normally long
// finished and the chain wait is a single L2 read; a waiting CTA only ever waits
// on a smaller ticket, which is held by a running CTA, so the chain cannot
// deadlock.  Regions too large for the shared window, and regions that need (2),
// run the same code on the HBM slot itself (the window pointer is generic).
//
// Float arithmetic mirrors the reference expression by expression and the file
// is compiled with -fmad=false: PCM must be bit-exact.  The one place an FMA is
// used is the pitch pre-filter (estimate_pitch_pair), whose results only select
// which lags are then evaluated exactly.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "block_prims.cuh"
#include "ctts_plan.h"

namespace ctts {

constexpr int ASM_THREADS = 256;
constexpr int ASM_WARPS = ASM_THREADS / 32;
constexpr int PITCH_FRAME = 256;  // ctts.c:2194
constexpr int LUT_N = 1024;       // ctts.c:52
constexpr int CONTOUR_KPT = 4;    // outputs per thread per contour tile

// private op kind: an op the host proved to be a no-op (plan compile step)
constexpr uint16_t OP_NOP = 0;

struct DevTables {
    const float* fade_out;  // 1 -> 0 raised cosine
    const float* fade_in;   // 0 -> 1 raised cosine
    const float* sine;      // quarter sine
    const float* hann256;
    const float* hann512;
};

enum { TASK_LAST = 1u, TASK_TO_PRE = 2u, TASK_GLOBAL = 4u };

struct RegionTask {
    uint32_t utt;       // index into out_counts / pre_counts / err
    uint32_t op_begin;
    uint32_t op_end;
    uint32_t bound;     // upper bound of the samples this task appends
    int32_t pred;       // task of the same utterance that precedes this one, -1: none
    uint32_t flags;     // TASK_*
    uint32_t dst_cap;   // utterance slot capacity in samples
    uint32_t big;       // slot in the global trim scratch, 0xffffffff: none
    unsigned long long dst_off;  // sample offset of the utterance slot in dst
};

struct AsmArgs {
    const int16_t* pool;        // re-packed PCM pool, every unit 16-byte aligned, zero padded to 8
    const uint32_t* unit_off;   // samples, multiple of 8
    const uint32_t* unit_cnt;
    uint32_t n_units;
    DevTables tab;
    const ctts_plan_op* ops;
    const RegionTask* tasks;    // in ticket order
    uint32_t n_tasks;
    int16_t* dst_final;
    int16_t* dst_pre;
    uint32_t* out_counts;
    uint32_t* pre_counts;
    uint32_t* err;              // per utterance, 0 = ok
    uint32_t* trim_scratch;     // global fallback for the silence bitmask
    uint32_t trim_scratch_words;  // per slot
    unsigned long long* chain;  // per task: (epoch << 32) | inclusive sample count
    uint32_t* ticket;           // zeroed before every launch
    uint32_t epoch;             // != 0, changes every launch
    ctts_assembly_params prm;
    uint32_t wcap;       // window capacity (samples, multiple of 8)
    uint32_t hcap;       // unit-head staging capacity (samples, multiple of 8)
    uint32_t scr_words;  // shared scratch, 32-bit words
};

enum { ERR_WINDOW_OVERFLOW = 1, ERR_UNIT_TOO_LONG = 2, ERR_BAD_OP = 3, ERR_SLOT_OVERFLOW = 4 };

// float -> int16 as x86-64 gcc compiles `(int16_t)f`: cvttss2si, keep low 16 bits
__device__ __forceinline__ int16_t f2s(float v) { return (int16_t)(int32_t)v; }

__device__ __forceinline__ float clamp16f(float v) {
    if (v > 32767.0f) v = 32767.0f;
    if (v < -32768.0f) v = -32768.0f;
    return v;
}

// fast_fade_out / fast_fade_in / fast_sine_fade, ctts.c:76-101
__device__ __forceinline__ float lut_lerp(const float* __restrict__ lut, float t) {
    float x = t * (float)(LUT_N - 1);
    int k = (int)x;
    if (k >= LUT_N - 1) return __ldg(lut + LUT_N - 1);
    if (k < 0) return __ldg(lut);
    float fr = x - (float)k;
    return __ldg(lut + k) * (1.0f - fr) + __ldg(lut + k + 1) * fr;
}

// abs() the way the reference computes it on int16 (ctts.c:1641): -32768 stays -32768
__device__ __forceinline__ int abs16(int16_t v) { return (int)(int16_t)(v > 0 ? v : -v); }

struct Smem {
    int16_t* win;                // wcap + 16 samples
    int16_t* hstage;             // hcap samples: the head of the unit being joined
    uint32_t* scratch;           // scr_words
    float* hann256;
    float* nrm2;                 // hann256[i+128] + hann256[i], 128 entries
    unsigned long long* red;     // 2 * ASM_WARPS entries
    uint32_t* bcast;             // 4 words
};

// Per-CTA execution state (replicated in every thread; all control flow is CTA-uniform).
// Sample indices are relative to the first sample of the task; in HBM mode w = dst + base,
// so negative indices reach the finished samples of earlier tasks.
struct State {
    int16_t* w;          // the window: w[i], i in [in_smem ? 0 : -base, cap)
    uint32_t cap;
    bool in_smem;
    bool have_base;
    uint32_t base;       // absolute sample count at the start of the task (valid iff have_base)
    uint32_t cnt;        // samples appended by this task so far: buf.count == base + cnt
    uint32_t word_start; // word_start_sample - base
    int32_t pred;
    int16_t* dst;        // utterance slot in HBM
    uint32_t dst_cap;
    uint32_t err;
};

// ---------------------------------------------------------------- look-back chain

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Block until the predecessor task has published its inclusive count.
__device__ void need_base(State& s, const Smem& sm, const AsmArgs& A) {
    if (s.have_base) return;
    if (threadIdx.x == 0) {
        const unsigned long long* p = A.chain + s.pred;
        unsigned long long v;
        unsigned ns = 20;
        while ((uint32_t)((v = ld_acquire_u64(p)) >> 32) != A.epoch) {
            __nanosleep(ns);
            if (ns < 640) ns *= 2;
        }
        sm.bcast[0] = (uint32_t)v;
    }
    __syncthreads();
    s.base = sm.bcast[0];
    s.have_base = true;
    __syncthreads();
}

// ---------------------------------------------------------------- window moves

// dst[base + a .. base + b) <- win[a..b): the source is 2-byte aligned only (base is arbitrary),
// the destination is written with 16-byte stores.
__device__ void flush_window(const State& s, const Smem& sm, uint32_t a, uint32_t b) {
    const int tid = threadIdx.x;
    if (b <= a) return;
    int16_t* d = s.dst + s.base;   // d[i] <-> win[i]
    const uint32_t phase = (uint32_t)((reinterpret_cast<uintptr_t>(d + a) >> 1) & 7u);
    uint32_t h = (8u - phase) & 7u;          // scalar head up to the first aligned vector
    if (h > b - a) h = b - a;
    if ((uint32_t)tid < h) d[a + tid] = sm.win[a + tid];
    const uint32_t v0 = a + h;               // first sample of vector 0
    const uint32_t nvec = (b - v0) >> 3;
    const uint32_t* w32 = reinterpret_cast<const uint32_t*>(sm.win);
    const uint32_t odd = v0 & 1u;
    const uint32_t bits = odd * 16u;
    int4* dv = reinterpret_cast<int4*>(d + v0);
    if ((v0 & 7u) == 0) {
        const int4* sv = reinterpret_cast<const int4*>(sm.win + v0);
        for (uint32_t v = tid; v < nvec; v += ASM_THREADS) dv[v] = sv[v];
    } else {
        for (uint32_t v = tid; v < nvec; v += ASM_THREADS) {
            const uint32_t wi = (v0 + 8u * v) >> 1;
            uint32_t r0 = w32[wi], r1 = w32[wi + 1], r2 = w32[wi + 2], r3 = w32[wi + 3], r4 = w32[wi + 4];
            int4 q;
            q.x = (int)__funnelshift_r(r0, r1, bits);
            q.y = (int)__funnelshift_r(r1, r2, bits);
            q.z = (int)__funnelshift_r(r2, r3, bits);
            q.w = (int)__funnelshift_r(r3, r4, bits);
            dv[v] = q;
        }
    }
    const uint32_t t0 = v0 + (nvec << 3);
    if (t0 + tid < b) d[t0 + tid] = sm.win[t0 + tid];
}

// Continue this task on the HBM slot (needs the base): the window becomes dst + base.
__device__ void enter_global(State& s, const Smem& sm, const AsmArgs& A) {
    need_base(s, sm, A);
    if (s.in_smem) {
        __syncthreads();
        flush_window(s, sm, 0, s.cnt);
        s.in_smem = false;
        s.w = s.dst + s.base;
        s.cap = s.dst_cap > s.base ? s.dst_cap - s.base : 0u;
    }
    // make the predecessors' finished samples (and our own flush) visible to every thread
    __threadfence();
    __syncthreads();
}

// ---------------------------------------------------------------- pitch

// estimate_pitch (ctts.c:1899) for two signals of the same length at once (the
// buffer tail `a` and the unit head `b`).
//
// The reference evaluates, for each lag in 55..275, three sequential float sums
// over i < 220: corr += s[i]*s[i+lag], e1 += s[i]^2, e2 += s[i+lag]^2, and keeps
// the first lag whose corr/sqrtf(e1*e2) is the strict maximum (voiced iff > 0.3).
// Those sums cannot be reordered, and at 3 non-fused FP32 operations per
// (lag, i) pair they are 40 % of all instructions of the assembly path.  So
// the search is done in two steps that together give the identical result:
//
//  1. FILTER: every lag gets an approximate score a[lag] = c~ / sqrtf(e1x * e2x),
//     c~ accumulated with FMA (one instruction per pair, four lags per thread
//     sharing the operand loads) and e1x, e2x EXACT integer window sums taken
//     from a 64-bit prefix sum of the squares.  For 220 terms |a - r| <= 4.2e-5
//     where r is the reference's score (standard summation error bound,
//     n*u*sum|x_i*y_i| <= n*u*sqrt(e1*e2) by Cauchy-Schwarz; DESIGN.md derives it).
//  2. EXACT: with eps = 1e-3 (24 x the bound), a signal is unvoiced if
//     max a <= 0.3 - eps; otherwise only lags with a >= max a - 2*eps can be the
//     reference's arg max, and those (typically 1-3) are evaluated by one thread
//     each with the reference's exact operation order.
//
// Both signals are needed voiced by the caller, so step 2 is skipped entirely
// when either signal fails the filter.
constexpr int PITCH_LO = CTTS_PLAN_SAMPLE_RATE / 400;  // 55
constexpr int PITCH_HI = CTTS_PLAN_SAMPLE_RATE / 80;   // 275
constexpr int PITCH_LEN = CTTS_PLAN_SAMPLE_RATE / 100; // 220
constexpr int PITCH_LAG0 = 53;        // lag of thread 0 (= 1 mod 4 keeps both float4 loads aligned)
constexpr int PITCH_LPT = 4;          // lags per thread
constexpr int PITCH_TPS = 64;         // threads per signal (57 used)
constexpr int PITCH_Y = 512;          // staged floats per signal (zero padded)
constexpr int PITCH_S = 504;          // prefix entries per signal
constexpr int PITCH_MAX_CAND = 64;    // per signal, beyond that: every lag is evaluated exactly
constexpr float PITCH_EPS = 1e-3f;
constexpr int PITCH_SCRATCH_WORDS = 2 * PITCH_Y + 2 * 2 * PITCH_S + 4 + 2 * (PITCH_MAX_CAND + 2) + 16;
static_assert(PITCH_LAG0 % 4 == 1 && PITCH_LAG0 <= PITCH_LO, "lag tiling");
static_assert(PITCH_LAG0 + PITCH_LPT * 57 > PITCH_HI, "57 threads cover every lag");
static_assert(PITCH_HI + PITCH_LPT + PITCH_LEN + 8 <= PITCH_Y, "staging covers the loop's reads");
static_assert(PITCH_HI + PITCH_LEN < PITCH_S, "prefix covers every window");

// exact score of one lag in the reference's order (ctts.c:1917-1931); lag 0 yields e1 in *e2_out
__device__ __forceinline__ float pitch_exact_sums(const float* y, uint32_t lag, uint32_t len, float* e2_out) {
    float c = 0.0f, e2 = 0.0f;
    const float* x = y;
    const float* z = y + lag;
#pragma unroll 4
    for (uint32_t i = 0; i < len; i++) {
        const float a = x[i], b = z[i];
        c += a * b;
        e2 += b * b;
    }
    *e2_out = e2;
    return c;
}

__device__ void estimate_pitch_pair(const Smem& sm, const int16_t* a, const int16_t* b, uint32_t n,
                                    float* pa, float* pb) {
    *pa = 0.0f;
    *pb = 0.0f;
    if (n < 200) return;
    const int tid = threadIdx.x, lane = lane_id(), warp = warp_id();
    const uint32_t lo = PITCH_LO;
    uint32_t hi = PITCH_HI;
    if (hi > n / 2) hi = n / 2;
    uint32_t len = PITCH_LEN;
    if (len > n - hi) len = n - hi;
    const uint32_t need = len + hi;  // <= 495 samples of each signal are ever read

    float* ya = reinterpret_cast<float*>(sm.scratch);
    float* yb = ya + PITCH_Y;
    unsigned long long* Sa = reinterpret_cast<unsigned long long*>(yb + PITCH_Y);  // 8-byte aligned: 2*PITCH_Y even
    unsigned long long* Sb = Sa + PITCH_S;
    unsigned long long* keys = Sb + PITCH_S;                      // [2]
    uint32_t* cand = reinterpret_cast<uint32_t*>(keys + 2);       // [2][PITCH_MAX_CAND + 2]
    uint32_t* ncand = cand + 2 * (PITCH_MAX_CAND + 2);            // [2]
    float* amax = reinterpret_cast<float*>(ncand + 2);            // [4] per lag warp
    float* e1s = amax + 4;                                        // [2]

    for (uint32_t i = tid; i < PITCH_Y; i += ASM_THREADS) {
        const bool in = i < need;
        ya[i] = in ? (float)a[i] : 0.0f;
        yb[i] = in ? (float)b[i] : 0.0f;
    }
    if (tid < 2) {
        ncand[tid] = 0;
        keys[tid] = 0ull;
    }
    __syncthreads();

    // ---- step 1: FMA scores on warps 0-3, exact prefix sums of squares on warps 4-5
    float c[PITCH_LPT] = {0.0f, 0.0f, 0.0f, 0.0f};
    const int sig = (tid >> 6) & 1;
    const uint32_t lag0 = PITCH_LAG0 + PITCH_LPT * (uint32_t)(tid & (PITCH_TPS - 1));
    const bool lag_thread = tid < 2 * PITCH_TPS && lag0 <= hi;
    if (lag_thread) {
        const float* x = sig ? yb : ya;
        const float* y = x + lag0;  // y[j] = s[lag0 + j]; (lag0 + 3) % 4 == 0
        float w0 = y[0], w1 = y[1], w2 = y[2];
        const uint32_t len4 = len & ~3u;
        for (uint32_t i = 0; i < len4; i += 4) {
            const float4 xv = *reinterpret_cast<const float4*>(x + i);
            const float4 yn = *reinterpret_cast<const float4*>(y + i + 3);
            const float w3 = yn.x, w4 = yn.y, w5 = yn.z, w6 = yn.w;
            c[0] = __fmaf_rn(xv.x, w0, c[0]); c[1] = __fmaf_rn(xv.x, w1, c[1]);
            c[2] = __fmaf_rn(xv.x, w2, c[2]); c[3] = __fmaf_rn(xv.x, w3, c[3]);
            c[0] = __fmaf_rn(xv.y, w1, c[0]); c[1] = __fmaf_rn(xv.y, w2, c[1]);
            c[2] = __fmaf_rn(xv.y, w3, c[2]); c[3] = __fmaf_rn(xv.y, w4, c[3]);
            c[0] = __fmaf_rn(xv.z, w2, c[0]); c[1] = __fmaf_rn(xv.z, w3, c[1]);
            c[2] = __fmaf_rn(xv.z, w4, c[2]); c[3] = __fmaf_rn(xv.z, w5, c[3]);
            c[0] = __fmaf_rn(xv.w, w3, c[0]); c[1] = __fmaf_rn(xv.w, w4, c[1]);
            c[2] = __fmaf_rn(xv.w, w5, c[2]); c[3] = __fmaf_rn(xv.w, w6, c[3]);
            w0 = w4; w1 = w5; w2 = w6;
        }
        for (uint32_t i = len4; i < len; i++) {
            const float xs = x[i];
#pragma unroll
            for (int k = 0; k < PITCH_LPT; k++) c[k] = __fmaf_rn(xs, y[i + k], c[k]);
        }
    } else if (warp == 4 || warp == 5) {
        // S[i] = sum_{j<i} s[j]^2, exact (values are int16, 504 * 2^30 < 2^64)
        const float* y = warp == 5 ? yb : ya;
        unsigned long long* S = warp == 5 ? Sb : Sa;
        constexpr int PER = 16;  // 32 lanes * 16 = 512 >= PITCH_S
        unsigned long long loc = 0;
        const int i0 = lane * PER;
#pragma unroll
        for (int k = 0; k < PER; k++) {
            const int v = (int)y[i0 + k];
            loc += (unsigned long long)(uint32_t)(v * v);
        }
        unsigned long long inc = loc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        unsigned long long run = inc - loc;  // exclusive
#pragma unroll
        for (int k = 0; k < PER; k++) {
            if (i0 + k < PITCH_S) S[i0 + k] = run;
            const int v = (int)y[i0 + k];
            run += (unsigned long long)(uint32_t)(v * v);
        }
    }
    __syncthreads();

    // ---- scores and per-signal maximum
    float sc[PITCH_LPT];
    float my_max = -1.0f;
    if (lag_thread) {
        const unsigned long long* S = sig ? Sb : Sa;
        const float e1 = (float)(S[len] - S[0]);
#pragma unroll
        for (int k = 0; k < PITCH_LPT; k++) {
            const uint32_t lag = lag0 + k;
            sc[k] = -1.0f;
            if (lag >= lo && lag <= hi) {
                const float e2 = (float)(S[lag + len] - S[lag]);
                const float nrm = sqrtf(e1 * e2);
                sc[k] = nrm > 0.0f ? c[k] / nrm : 0.0f;
                my_max = fmaxf(my_max, sc[k]);
            }
        }
    }
    if (tid < 2 * PITCH_TPS) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) my_max = fmaxf(my_max, __shfl_xor_sync(0xffffffffu, my_max, o));
        if (lane == 0) amax[warp] = my_max;
    }
    __syncthreads();
    const float max_a = fmaxf(amax[0], amax[1]), max_b = fmaxf(amax[2], amax[3]);
    // unvoiced by the filter: the reference's best cannot exceed 0.3
    if (!(max_a > 0.3f - PITCH_EPS) || !(max_b > 0.3f - PITCH_EPS)) {   // CTA-uniform
        __syncthreads();   // scratch is reused by the caller
        return;
    }

    // ---- candidates
    if (lag_thread) {
        const float thr = (sig ? max_b : max_a) - 2.0f * PITCH_EPS;
#pragma unroll
        for (int k = 0; k < PITCH_LPT; k++) {
            if (sc[k] >= thr) {
                uint32_t slot = atomicAdd(ncand + sig, 1u);
                if (slot < PITCH_MAX_CAND) cand[sig * (PITCH_MAX_CAND + 2) + slot] = lag0 + k;
            }
        }
    }
    __syncthreads();
    uint32_t na = ncand[0], nb = ncand[1];
    const bool all_a = na > PITCH_MAX_CAND, all_b = nb > PITCH_MAX_CAND;  // degenerate: evaluate every lag
    if (all_a) na = hi - lo + 1;
    if (all_b) nb = hi - lo + 1;
    __syncthreads();

    // ---- step 2: exact evaluation, one thread per (signal, lag); job 0 of each signal is lag 0 (= e1)
    const uint32_t jobs = na + nb + 2;
    for (uint32_t j = tid; j < jobs; j += ASM_THREADS) {
        const int sg = j < na + 1 ? 0 : 1;
        const uint32_t jj = sg ? j - (na + 1) : j;
        if (jj == 0) {
            float e1;
            (void)pitch_exact_sums(sg ? yb : ya, 0, len, &e1);
            e1s[sg] = e1;
        } else {
            const bool all = sg ? all_b : all_a;
            const uint32_t lag = all ? lo + (jj - 1) : cand[sg * (PITCH_MAX_CAND + 2) + (jj - 1)];
            float e2;
            const float cc = pitch_exact_sums(sg ? yb : ya, lag, len, &e2);
            // park the raw sums; the score needs e1, which another thread is computing
            reinterpret_cast<float2*>(sg ? Sb : Sa)[jj] = make_float2(cc, e2);   // S is dead from here on
        }
    }
    __syncthreads();
    for (uint32_t j = tid; j < jobs; j += ASM_THREADS) {
        const int sg = j < na + 1 ? 0 : 1;
        const uint32_t jj = sg ? j - (na + 1) : j;
        if (jj == 0) continue;
        const bool all = sg ? all_b : all_a;
        const uint32_t lag = all ? lo + (jj - 1) : cand[sg * (PITCH_MAX_CAND + 2) + (jj - 1)];
        const float2 ce = reinterpret_cast<const float2*>(sg ? Sb : Sa)[jj];
        float v = ce.x;
        const float nrm = sqrtf(e1s[sg] * ce.y);
        if (nrm > 0) v /= nrm;
        // the reference keeps the first lag that is strictly greater than everything before it,
        // starting from 0: the maximum positive score, smallest lag on ties
        if (v > 0.0f) atomicMax(keys + sg, ((unsigned long long)__float_as_uint(v) << 32) | (0xffffffffu - lag));
    }
    __syncthreads();
    const unsigned long long ka = keys[0], kb = keys[1];
    {
        const float v = __uint_as_float((uint32_t)(ka >> 32));
        const uint32_t l = 0xffffffffu - (uint32_t)(ka & 0xffffffffu);
        if (ka != 0ull && v > 0.3f && l > 0) *pa = (float)CTTS_PLAN_SAMPLE_RATE / (float)l;
    }
    {
        const float v = __uint_as_float((uint32_t)(kb >> 32));
        const uint32_t l = 0xffffffffu - (uint32_t)(kb & 0xffffffffu);
        if (kb != 0ull && v > 0.3f && l > 0) *pb = (float)CTTS_PLAN_SAMPLE_RATE / (float)l;
    }
    __syncthreads();   // scratch is reused by the caller
}

// smooth_pitch_boundary + apply_pitch_shift, ctts.c:1946-2024.  `reg` is the analysis length
// min(2*xf, count/2, n/2) resolved by the caller (0 = the reference returns early); `us` is the
// staged unit head.
__device__ void smooth_pitch(const State& s, const Smem& sm, int16_t* us, uint32_t n, uint32_t xf, uint32_t reg) {
    if (reg == 0) return;
    const int tid = threadIdx.x;
    float pp, np;
    estimate_pitch_pair(sm, s.w + ((int)s.cnt - (int)reg), us, reg, &pp, &np);
    if (!(pp > 0 && np > 0)) return;
    float ratio = np / pp;
    if (!(ratio > 1.15f || ratio < 0.85f)) return;
    float target = (ratio > 1.0f) ? 1.0f + (ratio - 1.0f) * 0.5f : 1.0f - (1.0f - ratio) * 0.5f;
    float shift = target / ratio;
    uint32_t len = xf;
    if (len > n / 4) len = n / 4;
    int16_t* tmp = reinterpret_cast<int16_t*>(sm.scratch);  // len <= hcap <= 2 * scr_words
    bool do_shift = !(shift < 0.9f || shift > 1.1f || len < 100);
    uint32_t keep = len;
    if (do_shift) {
        uint32_t m = (uint32_t)(unsigned long long)((float)len / shift);
        keep = m < len ? m : len;
    }
    for (uint32_t i = tid; i < len; i += ASM_THREADS) {
        int16_t r = 0;
        if (!do_shift) {
            r = us[i];
        } else if (i < keep) {
            float x = (float)i * shift;
            uint32_t k = (uint32_t)(unsigned long long)x;
            float fr = x - (float)k;
            if (k + 1 < len) r = f2s((float)us[k] * (1.0f - fr) + (float)us[k + 1] * fr);
            else if (k < len) r = us[k];
        }
        tmp[i] = r;
    }
    __syncthreads();
    for (uint32_t i = tid; i < len; i += ASM_THREADS) {
        float t = (float)i / (float)len;
        us[i] = f2s((float)tmp[i] * (1.0f - t) + (float)us[i] * t);
    }
    __syncthreads();
}

// match_boundary_energy, ctts.c:1730 (sums of squares are exact integers); len = min(xf, count, n)
__device__ void match_energy(const State& s, const Smem& sm, int16_t* us, uint32_t len) {
    if (len == 0) return;
    const int tid = threadIdx.x;
    const int16_t* tail = s.w + ((int)s.cnt - (int)len);
    long long sp = 0, sn = 0;
    for (uint32_t i = tid; i < len; i += ASM_THREADS) {
        int p = tail[i], q = us[i];
        sp += (long long)p * p;
        sn += (long long)q * q;
    }
    block_allreduce_add2<ASM_THREADS>(sp, sn, reinterpret_cast<long long*>(sm.red));
    float pr = (float)sqrt((double)sp / (double)len);
    float nr = (float)sqrt((double)sn / (double)len);
    if (pr < 1.0f || nr < 1.0f) return;
    float ratio = pr / nr;
    if (ratio > 2.0f) ratio = 2.0f;
    if (ratio < 0.5f) ratio = 0.5f;
    const float flen = (float)len;
    for (uint32_t i = tid; i < len; i += ASM_THREADS) {
        float t = (float)i / flen;
        float g = ratio * (1.0f - t) + 1.0f * t;
        us[i] = f2s(clamp16f((float)us[i] * g));
    }
    __syncthreads();
}

// ---------------------------------------------------------------- unit op

// the two crossfade gains at t share the table position (fast_fade_out / fast_fade_in, ctts.c:76-92)
__device__ __forceinline__ void crossfade_gains(const DevTables& tab, float t, float* pg, float* ng) {
    float x = t * (float)(LUT_N - 1);
    int k = (int)x;
    if (k >= LUT_N - 1) {
        *pg = __ldg(tab.fade_out + LUT_N - 1);
        *ng = __ldg(tab.fade_in + LUT_N - 1);
    } else if (k < 0) {
        *pg = __ldg(tab.fade_out);
        *ng = __ldg(tab.fade_in);
    } else {
        float fr = x - (float)k, om = 1.0f - fr;
        *pg = __ldg(tab.fade_out + k) * om + __ldg(tab.fade_out + k + 1) * fr;
        *ng = __ldg(tab.fade_in + k) * om + __ldg(tab.fade_in + k + 1) * fr;
    }
}

__device__ __forceinline__ int sub_dc(int v, int dc) {
    // clamp(v - dc) to int16 (remove_dc_offset, ctts.c:1577-1581)
    return max(__viaddmin_s32(v, -dc, 32767), -32768);
}

// normalize_rms's per-sample step (ctts.c:1720-1725)
__device__ __forceinline__ int scale_sample(int x, bool scale, float g) {
    return scale ? (int)f2s(clamp16f((float)x * g)) : x;
}

// 8 consecutive samples starting `sh` samples (1..8) into the 16-sample pair (lo, hi)
__device__ __forceinline__ int4 shift_pick(const int4& lo, const int4& hi, uint32_t sh) {
    const uint32_t r0 = lo.x, r1 = lo.y, r2 = lo.z, r3 = lo.w, r4 = hi.x, r5 = hi.y, r6 = hi.z, r7 = hi.w;
    uint32_t a0, a1, a2, a3, a4;
    switch (sh >> 1) {   // CTA-uniform
        case 0: a0 = r0; a1 = r1; a2 = r2; a3 = r3; a4 = r4; break;
        case 1: a0 = r1; a1 = r2; a2 = r3; a3 = r4; a4 = r5; break;
        case 2: a0 = r2; a1 = r3; a2 = r4; a3 = r5; a4 = r6; break;
        case 3: a0 = r3; a1 = r4; a2 = r5; a3 = r6; a4 = r7; break;
        default: a0 = r4; a1 = r5; a2 = r6; a3 = r7; a4 = 0u; break;
    }
    const uint32_t bits = (sh & 1u) * 16u;
    int4 q;
    q.x = (int)__funnelshift_r(a0, a1, bits);
    q.y = (int)__funnelshift_r(a1, a2, bits);
    q.z = (int)__funnelshift_r(a2, a3, bits);
    q.w = (int)__funnelshift_r(a3, a4, bits);
    return q;
}

// ctts.c:3785-3846: gather -> normalize_rms -> [smooth, match] -> buffer_append_crossfade.
//
// The unit's first `hs` samples (everything the join may rewrite, and what the pitch analysis
// reads) are staged in `hstage`; the rest is written straight to its final place in the window,
// on the window's own 16-byte grid (the pool side is re-aligned with a funnel shift), and
// revisited once in place to subtract the DC offset -- which is only known after the head has
// been smoothed and energy matched.
__device__ void op_unit(State& s, const Smem& sm, const AsmArgs& A, const ctts_plan_op& op) {
    const int tid = threadIdx.x;
    if (op.a >= A.n_units) { s.err = ERR_BAD_OP; return; }
    const uint32_t n = __ldg(A.unit_cnt + op.a);
    if (n == 0) return;
    const int16_t* src = A.pool + __ldg(A.unit_off + op.a);
    const int4* srcv = reinterpret_cast<const int4*>(src);
    const uint32_t nvec = (n + 7) >> 3;
    int16_t* us = sm.hstage;
    const uint32_t xf = op.b;
    const bool boundary = (op.flags & CTTS_UNIT_AFTER_BOUNDARY) != 0;
    const bool remove_dc = A.prm.remove_dc_offset != 0;

    // ---- decisions that depend on buf.count = base + cnt; the base is only waited for when
    //      the samples of this task alone cannot settle them
    bool join = false;
    if (!boundary) {
        if (s.cnt == 0) need_base(s, sm, A);
        join = s.cnt > 0 || s.base > 0;
    }
    // !join <=> count == 0 || after_word_boundary: the unit starts fresh (fade-in, no crossfade)
    uint32_t a = 0;             // crossfade = energy-match length min(xf, count, n), ctts.c:3319, :1736
    uint32_t reg = 0;           // pitch analysis length, ctts.c:1983-1987
    if (join && xf > 0) {
        const uint32_t m = xf < n ? xf : n;
        if (s.cnt >= m) a = m;
        else {
            need_base(s, sm, A);
            const unsigned long long count = (unsigned long long)s.base + s.cnt;
            a = count < m ? (uint32_t)count : m;
        }
        if (n >= 200) {
            const uint32_t m2 = 2 * xf < n / 2 ? 2 * xf : n / 2;
            if (s.cnt >= 200 && s.cnt / 2 >= m2) reg = m2;
            else {
                need_base(s, sm, A);
                const unsigned long long count = (unsigned long long)s.base + s.cnt;
                if (count >= 200) reg = count / 2 < m2 ? (uint32_t)(count / 2) : m2;
            }
        }
        // a window that reaches back past the start of this task: continue on the HBM slot
        if ((a > s.cnt || reg > s.cnt) && s.in_smem) enter_global(s, sm, A);
    }
    if ((unsigned long long)s.cnt + (n - a) > s.cap) { s.err = ERR_WINDOW_OVERFLOW; return; }

    // staged head: what smooth/match may rewrite (min(xf, n)) and what the pitch analysis reads (<= 495)
    uint32_t hs = 0;
    if (join) {
        uint32_t want = xf < n ? xf : n;
        if (reg > 0 && want < 496) want = 496;
        hs = (want + 7) & ~7u;
        if (hs > (nvec << 3)) hs = nvec << 3;
        if (hs > A.hcap) { s.err = ERR_UNIT_TOO_LONG; return; }
    }
    const uint32_t hsn = hs < n ? hs : n;   // staged samples that exist

    // ---- pass 1: sum of squares (normalize_rms, ctts.c:1709; double sum of integers == integer sum)
    long long ss = 0;
    for (uint32_t v = tid; v < nvec; v += ASM_THREADS) {
        const int4 q = __ldg(srcv + v);
        const int16_t* e = reinterpret_cast<const int16_t*>(&q);
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int x = e[k];
            ss += (long long)(x * x);
        }
    }
    ss = block_allreduce<ASM_THREADS>(ss, OpAddI64(), reinterpret_cast<long long*>(sm.red));
    bool scale = false;
    float g = 1.0f;
    if (A.prm.target_rms > 0) {
        float rms = (float)sqrt((double)ss / (double)n);
        if (!(rms < 1.0f)) {
            g = A.prm.target_rms / rms;
            if (g > 3.0f) g = 3.0f;
            if (g < 0.1f) g = 0.1f;
            scale = true;
        }
    }

    // ---- pass 2: scale; head -> hstage, body -> window (aligned vectors of the window)
    for (uint32_t v = tid; v < (hs >> 3); v += ASM_THREADS) {
        int4 q = __ldg(srcv + v);
        int16_t* e = reinterpret_cast<int16_t*>(&q);
#pragma unroll
        for (int k = 0; k < 8; k++) e[k] = (int16_t)scale_sample((int)e[k], scale, g);
        *(reinterpret_cast<int4*>(us) + v) = q;
    }
    // unit sample i lands at tail[i]; the body is [hs, n)
    int16_t* tail = s.w + ((int)s.cnt - (int)a);
    const uint32_t phase = (uint32_t)((reinterpret_cast<uintptr_t>(tail + hs) >> 1) & 7u);
    int16_t* grid = tail + hs - phase;                       // 16-byte aligned
    const uint32_t body = n - hsn;                            // may be 0
    const uint32_t gvec = body ? (phase + body + 7) >> 3 : 0; // window vectors that hold body samples
    const uint32_t pv0 = hs >> 3;                             // pool vector of unit sample hs
    int dsum = 0;
    for (uint32_t j = tid; j < gvec; j += ASM_THREADS) {
        // window vector j holds unit samples i0 .. i0+7, i0 = hs - phase + 8j
        int4 q;
        if (phase == 0) {
            q = __ldg(srcv + pv0 + j);
        } else {
            int4 lo = make_int4(0, 0, 0, 0), hi = make_int4(0, 0, 0, 0);
            if (pv0 + j >= 1) lo = __ldg(srcv + pv0 + j - 1);
            if (pv0 + j < nvec) hi = __ldg(srcv + pv0 + j);
            q = shift_pick(lo, hi, 8u - phase);
        }
        int16_t* e = reinterpret_cast<int16_t*>(&q);
        const int i0 = (int)hs - (int)phase + 8 * (int)j;
        const bool full = i0 >= (int)hs && i0 + 8 <= (int)n;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int y = scale_sample((int)e[k], scale, g);
            e[k] = (int16_t)y;
            if (full || (i0 + k >= (int)hs && i0 + k < (int)n)) dsum += y;
        }
        if (full) {
            *(reinterpret_cast<int4*>(grid) + j) = q;
        } else {
#pragma unroll
            for (int k = 0; k < 8; k++)
                if (i0 + k >= (int)hs && i0 + k < (int)n) grid[8 * j + k] = e[k];
        }
    }
    __syncthreads();

    if (join) {
        smooth_pitch(s, sm, us, n, xf, reg);
        match_energy(s, sm, us, a);
    }

    // ---- remove_dc_offset (ctts.c:1568) inside buffer_append_crossfade (ctts.c:3279)
    int dc = 0;
    if (remove_dc) {
        for (uint32_t i = tid; i < hsn; i += ASM_THREADS) dsum += us[i];
        long long sum = block_allreduce<ASM_THREADS>((long long)dsum, OpAddI64(), reinterpret_cast<long long*>(sm.red));
        dc = (int)(int16_t)(sum / (long long)n);
    }
    // body in place
    if (dc != 0) {
        for (uint32_t j = tid; j < gvec; j += ASM_THREADS) {
            const int i0 = (int)hs - (int)phase + 8 * (int)j;
            const bool full = i0 >= (int)hs && i0 + 8 <= (int)n;
            if (full) {
                int4 q = *(reinterpret_cast<int4*>(grid) + j);
                int16_t* e = reinterpret_cast<int16_t*>(&q);
#pragma unroll
                for (int k = 0; k < 8; k++) e[k] = (int16_t)sub_dc((int)e[k], dc);
                *(reinterpret_cast<int4*>(grid) + j) = q;
            } else {
#pragma unroll
                for (int k = 0; k < 8; k++)
                    if (i0 + k >= (int)hs && i0 + k < (int)n) grid[8 * j + k] = (int16_t)sub_dc((int)grid[8 * j + k], dc);
            }
        }
    }
    if (join) {
        // staged head: crossfade mix (ctts.c:3328-3344) over [0, a), plain copy over [a, hsn)
        const float inv = a ? 1.0f / (float)a : 0.0f;
        for (uint32_t i = tid; i < hsn; i += ASM_THREADS) {
            int v = us[i];
            if (remove_dc) v = sub_dc(v, dc);
            if (i < a) {
                float pg, ng;
                crossfade_gains(A.tab, (float)i * inv, &pg, &ng);
                int p = tail[i];
                int mix = (int)((float)p * pg + (float)v * ng);
                v = max(min(mix, 32767), -32768);
            }
            tail[i] = (int16_t)v;
        }
    } else {
        // fade-in of a word-initial unit (apply_fade_in, ctts.c:3015), after the DC removal
        const uint32_t pre = A.prm.fade_in_samples < n ? A.prm.fade_in_samples : n;
        if (pre) {
            __syncthreads();
            const float inv = 1.0f / (float)pre;
            for (uint32_t i = tid; i < pre; i += ASM_THREADS)
                tail[i] = f2s((float)tail[i] * lut_lerp(A.tab.sine, (float)i * inv));
        }
    }
    s.cnt += n - a;
    __syncthreads();
}

// ---------------------------------------------------------------- word end

// remove_silence_regions, ctts.c:1634, as a bitmask + scan + in-place compaction.
// Returns the new length.  `reg` = w + word_start, len = count - word_start.
__device__ uint32_t trim_region(const Smem& sm, const AsmArgs& A, uint32_t big, int16_t* reg, uint32_t len) {
    const int tid = threadIdx.x, lane = lane_id(), warp = warp_id();
    const uint32_t min_sil = A.prm.min_silence_samples;
    int pk = 0;
    for (uint32_t i = tid; i < len; i += ASM_THREADS) {
        int a = abs16(reg[i]);
        pk = a > pk ? a : pk;
    }
    pk = block_allreduce<ASM_THREADS>(pk, OpMaxI32(), reinterpret_cast<int*>(sm.red));
    if (pk == 0) return len;
    const int limit = (int)f2s((float)pk * A.prm.silence_threshold);
    uint32_t keep_n = min_sil / 4;
    if (keep_n < 10) keep_n = 10;

    const uint32_t wn = (len + 31) >> 5;
    uint32_t* words;
    if (2 * wn <= A.scr_words) words = sm.scratch;
    else words = A.trim_scratch + (size_t)big * A.trim_scratch_words;   // host sized it for this task
    uint32_t* woff = words + wn;

    // 1 bit per sample: |x| <= threshold
    for (uint32_t wd = warp; wd < wn; wd += ASM_THREADS / 32) {
        uint32_t i = (wd << 5) + lane;
        bool sil = (i < len) && (abs16(reg[i]) <= limit);
        uint32_t m = __ballot_sync(0xffffffffu, sil);
        if (lane == 0) words[wd] = m;
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
        uint32_t i0 = c0 + (uint32_t)tid * 8;
        int16_t v[8];
        uint32_t km = 0, d0 = 0;
        if (i0 < len) {
            uint32_t j = i0 >> 5, b = i0 & 31;  // 8 | 32: one word
            uint32_t kw = words[j];
            km = (kw >> b) & 0xffu;
            d0 = woff[j] + __popc(kw & ((1u << b) - 1u));
#pragma unroll
            for (int k = 0; k < 8; k++) v[k] = (i0 + k < len) ? reg[i0 + k] : (int16_t)0;
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
// tile by tile: the originals a tile needs ([t0-256, t1+288)) are staged in
// shared scratch (zero past the end of the segment: reads the reference performs
// past the end of its heap copy -- undefined behaviour there, DESIGN.md
// "Reference UB" -- yield 0 here), per-frame pitch factors come from a table, and
// the norm of an interior sample is the precomputed hann[i+128] + hann[i].
// When `energy` is set the linear energy ramp of apply_phrase_intonation
// (ctts.c:2857-2864) over the whole word (index ebase + j, denominator eden) is
// applied to each sample as it is written.  Returns false if nothing was done.
constexpr uint32_t CONTOUR_TILE = ASM_THREADS * CONTOUR_KPT;
constexpr uint32_t CONTOUR_STAGE = CONTOUR_TILE + 576;       // samples (256 behind, 288 ahead, 8 phase, pad)
constexpr uint32_t CONTOUR_PF_MAX = 1024;                    // frames with a tabulated pitch factor
constexpr uint32_t CONTOUR_SCRATCH_WORDS = CONTOUR_STAGE / 2 + CONTOUR_PF_MAX;

__device__ bool pitch_contour(const Smem& sm, int16_t* x, uint32_t n, float f0, float f1, bool energy,
                              float e0, float de, float eden, uint32_t ebase) {
    if (n < 100 || fabsf(f0 - f1) < 0.01f) return false;
    if (n < PITCH_FRAME) return false;  // no frame fits: every sample keeps its original value
    const int tid = threadIdx.x;
    const uint32_t frames = (n - PITCH_FRAME) / (PITCH_FRAME / 2) + 1;
    const bool degenerate = (n == PITCH_FRAME);  // 1/(n-256) = inf in the reference: NaN indices
    const float inv = 1.0f / (float)(n - PITCH_FRAME);
    const float* hann = sm.hann256;
    const float* nrm2 = sm.nrm2;
    int16_t* stage = reinterpret_cast<int16_t*>(sm.scratch);
    float* pft = reinterpret_cast<float*>(sm.scratch + CONTOUR_STAGE / 2);
    const bool tabulated = frames <= CONTOUR_PF_MAX;
    if (tabulated) {
        for (uint32_t k = tid; k < frames; k += ASM_THREADS) {
            float t = (float)(k << 7) * inv;
            float st = t * t * (3.0f - 2.0f * t);
            pft[k] = f0 + (f1 - f0) * st;
        }
    }
    // stage[phase + 256 + u] = x[t0 + u]: same 16-byte phase on both sides
    const uint32_t phase = (uint32_t)((reinterpret_cast<uintptr_t>(x) >> 1) & 7u);
    int16_t* sbase = stage + phase + PITCH_FRAME;   // index u relative to the tile start
    for (uint32_t t0 = 0; t0 < n; t0 += CONTOUR_TILE) {
        const uint32_t t1 = min(t0 + CONTOUR_TILE, n);
        // ---- staging: carry the 544 samples that overlap the previous tile, load the new ones
        if (t0 != 0) {
            for (uint32_t v = tid; v < 560 / 8 + 1; v += ASM_THREADS) {
                int4* d = reinterpret_cast<int4*>(stage) + v;
                *d = *(reinterpret_cast<const int4*>(stage + CONTOUR_TILE) + v);
            }
        }
        __syncthreads();
        {
            // u range to load: [u_lo, u_hi) as whole 16-byte vectors of the staging buffer
            const int u_lo = t0 == 0 ? -(int)phase : (int)(288 + 8 - phase) & ~7;  // first vector not carried
            const int first_vec = (int)(phase + PITCH_FRAME + u_lo) >> 3;
            const int last_vec = (int)(CONTOUR_STAGE >> 3);
            for (int v = first_vec + tid; v < last_vec; v += ASM_THREADS) {
                const int u0 = (v << 3) - (int)(phase + PITCH_FRAME);   // u of the vector's first sample
                const long long g0 = (long long)t0 + u0;                // segment index
                int4 q = make_int4(0, 0, 0, 0);
                if (g0 >= 0 && g0 + 8 <= (long long)n) {
                    q = *reinterpret_cast<const int4*>(x + g0);
                } else {
                    int16_t* e = reinterpret_cast<int16_t*>(&q);
#pragma unroll
                    for (int k = 0; k < 8; k++) {
                        long long g = g0 + k;
                        if (g >= 0 && g < (long long)n) e[k] = x[g];
                    }
                }
                *(reinterpret_cast<int4*>(stage) + v) = q;
            }
        }
        __syncthreads();
        // ---- outputs of this tile
#pragma unroll
        for (int r = 0; r < CONTOUR_KPT; r++) {
            const uint32_t j = t0 + (uint32_t)tid + (uint32_t)r * ASM_THREADS;
            if (j >= t1) continue;
            const uint32_t k1 = j >> 7;
            const uint32_t i1 = j & 127u;
            const bool v1 = k1 < frames;
            const bool v0 = k1 >= 1 && k1 - 1 < frames;
            int16_t acc = 0;
            float norm = 0.0f;
            if (v0) {
                const uint32_t kk = k1 - 1, i = i1 + 128u;
                float v = 0.0f;
                if (!degenerate) {
                    float pf;
                    if (tabulated) pf = pft[kk];
                    else {
                        float t = (float)(kk << 7) * inv;
                        pf = f0 + (f1 - f0) * (t * t * (3.0f - 2.0f * t));
                    }
                    float xs = (float)i * pf;
                    uint32_t k = (uint32_t)(unsigned long long)xs;
                    float fr = xs - (float)k;
                    const int16_t* sp = sbase + ((int)(kk << 7) - (int)t0) + (int)k;
                    v = (k + 1 < PITCH_FRAME) ? (float)sp[0] * (1.0f - fr) + (float)sp[1] * fr : (float)sp[0];
                }
                acc = f2s(v * hann[i]);
                norm = hann[i];
            }
            if (v1) {
                const uint32_t kk = k1, i = i1;
                float v = 0.0f;
                if (!degenerate) {
                    float pf;
                    if (tabulated) pf = pft[kk];
                    else {
                        float t = (float)(kk << 7) * inv;
                        pf = f0 + (f1 - f0) * (t * t * (3.0f - 2.0f * t));
                    }
                    float xs = (float)i * pf;
                    uint32_t k = (uint32_t)(unsigned long long)xs;
                    float fr = xs - (float)k;
                    const int16_t* sp = sbase + ((int)(kk << 7) - (int)t0) + (int)k;
                    v = (k + 1 < PITCH_FRAME) ? (float)sp[0] * (1.0f - fr) + (float)sp[1] * fr : (float)sp[0];
                }
                acc = (int16_t)(acc + f2s(v * hann[i]));
                norm = v0 ? nrm2[i1] : hann[i];
            }
            int16_t o;
            if (norm > 0.01f) o = f2s(clamp16f((float)acc / norm));
            else o = sbase[(int)(j - t0)];
            if (energy) {
                float t = (float)(j + ebase) / eden;
                o = f2s(clamp16f((float)o * (e0 + de * t)));
            }
            x[j] = o;
        }
        __syncthreads();
    }
    return true;
}

// ctts.c:3693-3713 / :3878-3898: trim then phrase intonation on [word_start, count)
__device__ void op_word_end(State& s, const Smem& sm, const AsmArgs& A, uint32_t big, const ctts_plan_op& op) {
    const int tid = threadIdx.x;
    if ((op.flags & CTTS_WE_TRIM) && s.cnt > s.word_start) {
        uint32_t len = s.cnt - s.word_start;
        if (len > A.prm.min_silence_samples)
            s.cnt = s.word_start + trim_region(sm, A, big, s.w + s.word_start, len);
    }
    if (s.cnt <= s.word_start) return;
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

// ---------------------------------------------------------------- kernel

// apply_fade_out on the buffer tail (ctts.c:3028): `if (buf.count > 0) apply_fade_out(buf, count, a)`
__device__ void op_fade_out(State& s, const Smem& sm, const AsmArgs& A, uint32_t a) {
    if (a == 0) return;
    uint32_t f = a;
    if (s.cnt < a) {
        need_base(s, sm, A);
        const unsigned long long count = (unsigned long long)s.base + s.cnt;
        if (count == 0) return;
        if (count < f) f = (uint32_t)count;
        if (f > s.cnt && s.in_smem) enter_global(s, sm, A);
    }
    int16_t* p = s.w + ((int)s.cnt - (int)f);
    const float inv = 1.0f / (float)f;
    for (uint32_t i = threadIdx.x; i < f; i += ASM_THREADS)
        p[i] = f2s((float)p[i] * lut_lerp(A.tab.sine, (float)(f - i) * inv));
    __syncthreads();
}

// buffer_append_silence, ctts.c:3361
__device__ void op_silence(State& s, uint32_t n) {
    if ((unsigned long long)s.cnt + n > s.cap) { s.err = ERR_WINDOW_OVERFLOW; return; }
    int16_t* p = s.w + s.cnt;
    const int tid = threadIdx.x;
    // scalar head to the 16-byte grid, vector body, scalar tail
    uint32_t h = (8u - (uint32_t)((reinterpret_cast<uintptr_t>(p) >> 1) & 7u)) & 7u;
    if (h > n) h = n;
    if ((uint32_t)tid < h) p[tid] = 0;
    const uint32_t nv = (n - h) >> 3;
    int4* pv = reinterpret_cast<int4*>(p + h);
    for (uint32_t v = tid; v < nv; v += ASM_THREADS) pv[v] = make_int4(0, 0, 0, 0);
    const uint32_t t0 = h + (nv << 3);
    if (t0 + tid < n) p[t0 + tid] = 0;
    s.cnt += n;
    __syncthreads();
}

__device__ void run_task(const Smem& sm, const AsmArgs& A, uint32_t ti) {
    const int tid = threadIdx.x;
    const RegionTask task = A.tasks[ti];
    State s;
    s.dst = ((task.flags & TASK_TO_PRE) ? A.dst_pre : A.dst_final) + task.dst_off;
    s.dst_cap = task.dst_cap;
    s.cnt = 0;
    s.word_start = 0;
    s.err = 0;
    s.pred = task.pred;
    s.have_base = task.pred < 0;
    s.base = 0;
    s.in_smem = true;
    s.w = sm.win;
    s.cap = A.wcap;
    if (task.flags & TASK_GLOBAL) {
        // the region does not fit the shared window: assemble it in place in the HBM slot
        need_base(s, sm, A);
        s.in_smem = false;
        s.w = s.dst + s.base;
        s.cap = s.dst_cap > s.base ? s.dst_cap - s.base : 0u;
        __threadfence();
        __syncthreads();
    }

    for (uint32_t k = task.op_begin; k < task.op_end && !s.err; k++) {
        ctts_plan_op op;
        {
            const int4* p = reinterpret_cast<const int4*>(A.ops + k);
            int4 lo = __ldg(p), hi = __ldg(p + 1);
            *reinterpret_cast<int4*>(&op) = lo;
            *(reinterpret_cast<int4*>(&op) + 1) = hi;
        }
        switch (op.kind) {
            case OP_NOP:
                break;
            case CTTS_OP_UNIT:
                op_unit(s, sm, A, op);
                break;
            case CTTS_OP_SILENCE:
                op_silence(s, op.a);
                break;
            case CTTS_OP_FADE_OUT:
                op_fade_out(s, sm, A, op.a);
                break;
            case CTTS_OP_WORD_END:
                op_word_end(s, sm, A, task.big, op);
                break;
            case CTTS_OP_MARK:   // word_start_sample = buf.count, ctts.c:3723, :3765
                s.word_start = s.cnt;
                break;
            default:
                s.err = ERR_BAD_OP;
        }
    }

    // ---- publish: the region's final position is base, known from the predecessor
    need_base(s, sm, A);
    __syncthreads();
    if (s.in_smem) {
        if ((unsigned long long)s.base + s.cnt > s.dst_cap) s.err = s.err ? s.err : ERR_SLOT_OVERFLOW;
        else flush_window(s, sm, 0, s.cnt);
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const uint32_t total = s.base + s.cnt;
        if (s.err) atomicMax(A.err + task.utt, s.err);
        if (task.flags & TASK_LAST) {
            if (task.flags & TASK_TO_PRE) A.pre_counts[task.utt] = total;
            else A.out_counts[task.utt] = total;
        }
        st_release_u64(A.chain + ti, ((unsigned long long)A.epoch << 32) | total);
    }
}

__global__ void __launch_bounds__(ASM_THREADS, 3) assemble_kernel(const AsmArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem sm;
    sm.win = reinterpret_cast<int16_t*>(smem_raw);
    sm.hstage = sm.win + A.wcap + 16;
    sm.scratch = reinterpret_cast<uint32_t*>(sm.hstage + A.hcap);
    sm.hann256 = reinterpret_cast<float*>(sm.scratch + A.scr_words);
    sm.nrm2 = sm.hann256 + PITCH_FRAME;
    sm.red = reinterpret_cast<unsigned long long*>(sm.nrm2 + PITCH_FRAME / 2);
    sm.bcast = reinterpret_cast<uint32_t*>(sm.red + 2 * ASM_WARPS);

    const int tid = threadIdx.x;
    for (int i = tid; i < PITCH_FRAME; i += ASM_THREADS) sm.hann256[i] = __ldg(A.tab.hann256 + i);
    // norm of a sample covered by two frames: (0 + w[i+128]) + w[i], the reference's accumulation order
    for (int i = tid; i < PITCH_FRAME / 2; i += ASM_THREADS)
        sm.nrm2[i] = __ldg(A.tab.hann256 + i + PITCH_FRAME / 2) + __ldg(A.tab.hann256 + i);
    __syncthreads();

    // persistent CTAs: tasks are taken in ticket order (see the header comment)
    for (;;) {
        if (tid == 0) sm.bcast[1] = atomicAdd(A.ticket, 1u);
        __syncthreads();
        const uint32_t ti = sm.bcast[1];
        __syncthreads();
        if (ti >= A.n_tasks) break;
        run_task(sm, A, ti);
    }
}

}  // namespace ctts
