// assemble.cuh -- the audio-assembly kernel: one CTA executes one utterance's
// plan ops in order, with CTA-wide parallelism inside every op.
//
// The reference is a sequential program over a growing buffer in which each
// join reads the processed tail of what came before (ctts.c:3835-3845), so
// the parallel axes are: utterances (one CTA each), and samples / lags / words
// of a bitmask inside an op.  The live tail of the utterance (a "window":
// the last HALO finished samples plus the region since the last word mark)
// is kept in shared memory; finished samples are streamed to HBM with 16-byte
// stores when a region closes.  Regions that cannot fit use the output slot in
// HBM as the window instead (same code: all ops address the window through an
// absolute-index generic pointer).
//
// Float arithmetic mirrors the reference expression by expression and the
// file is compiled with -fmad=false: PCM must be bit-exact.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "block_prims.cuh"
#include "ctts_plan.h"

namespace ctts {

constexpr int ASM_THREADS = 512;
constexpr int PITCH_FRAME = 256;  // ctts.c:2194
constexpr int LUT_N = 1024;       // ctts.c:52
constexpr int CONTOUR_KPT = 4;    // outputs per thread per contour tile

struct DevTables {
    const float* fade_out;  // 1 -> 0 raised cosine
    const float* fade_in;   // 0 -> 1 raised cosine
    const float* sine;      // quarter sine
    const float* hann256;
    const float* hann512;
};

struct UttTask {
    uint32_t utt;          // index into out_counts / pre_counts / err
    uint32_t op_begin;
    uint32_t op_end;
    uint32_t first_bound;  // upper bound of samples appended before the first MARK
    unsigned long long dst_off;  // sample offset of this utterance's slot in dst
    uint32_t dst_cap;      // slot capacity in samples
    uint32_t to_pre;       // 1: write the pre-stretch buffer (speed != 1), 0: final PCM
};

struct AsmArgs {
    const int16_t* pool;        // re-packed PCM pool, every unit 16-byte aligned
    const uint32_t* unit_off;   // samples, multiple of 8
    const uint32_t* unit_cnt;
    uint32_t n_units;
    DevTables tab;
    const ctts_plan_op* ops;    // MARK ops carry the next region's bound in .a
    const UttTask* tasks;
    uint32_t n_tasks;
    int16_t* dst_final;
    int16_t* dst_pre;
    uint32_t* out_counts;
    uint32_t* pre_counts;
    uint32_t* err;              // per utterance, 0 = ok
    uint32_t* trim_scratch;     // global fallback for the silence bitmask
    uint32_t trim_scratch_words;  // per utterance
    ctts_assembly_params prm;
    uint32_t wcap;   // window capacity (samples, multiple of 8)
    uint32_t ucap;   // unit staging capacity (samples, multiple of 8)
    uint32_t halo;   // finished samples kept writable behind the word mark (multiple of 8)
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
    int16_t* win;
    int16_t* ustage;
    uint32_t* scratch;
    float* hann256;
    float* nrm2;                 // hann256[i+128] + hann256[i], 128 entries
    unsigned long long* red;     // 2 * ASM_THREADS/32 entries
};

// Per-CTA execution state (replicated in every thread; all control flow is CTA-uniform).
struct State {
    int16_t* w;          // window addressed by ABSOLUTE sample index: w[abs]
    uint32_t base;       // first absolute index held by the window
    uint32_t cap;        // window capacity
    bool in_smem;
    uint32_t count;      // buf.count
    uint32_t word_start; // word_start_sample
    int16_t* dst;        // utterance slot in HBM (dst[abs])
    uint32_t dst_cap;
    uint32_t err;
};

// ---------------------------------------------------------------- window moves

// dst[a..b) <- w[a..b); a is a multiple of 8 and both sides are 16-byte aligned there.
__device__ __forceinline__ void flush_range(const State& s, uint32_t a, uint32_t b) {
    const int tid = threadIdx.x;
    uint32_t nvec = (b - a) >> 3;
    const int4* src = reinterpret_cast<const int4*>(s.w + a);
    int4* d = reinterpret_cast<int4*>(s.dst + a);
    for (uint32_t i = tid; i < nvec; i += ASM_THREADS) d[i] = src[i];
    for (uint32_t i = a + (nvec << 3) + tid; i < b; i += ASM_THREADS) s.dst[i] = s.w[i];
}

// Called at a word mark: decide where the next region lives, stream finished
// samples to HBM and slide the live tail to the front of the shared window.
__device__ void region_switch(State& s, const Smem& sm, const AsmArgs& A, uint32_t next_bound) {
    const int tid = threadIdx.x;
    uint32_t keep_from = s.count > A.halo ? ((s.count - A.halo) & ~7u) : 0u;
    if (keep_from < s.base) keep_from = s.base;
    bool want_smem = (unsigned long long)(s.count - keep_from) + next_bound + 8ull <= A.wcap;
    __syncthreads();
    if (s.in_smem) {
        if (want_smem) {
            flush_range(s, s.base, keep_from);
            uint32_t delta = keep_from - s.base;
            if (delta) {
                // slide [keep_from, count) down by delta (both multiples of 8)
                uint32_t n = s.count - keep_from;
                int16_t* win = sm.win;
                for (uint32_t c0 = 0; c0 < n; c0 += ASM_THREADS * 8) {
                    uint32_t i = c0 + tid * 8;
                    int4 v = make_int4(0, 0, 0, 0);
                    if (i < n) v = *reinterpret_cast<const int4*>(win + delta + i);
                    __syncthreads();
                    if (i < n) *reinterpret_cast<int4*>(win + i) = v;
                    __syncthreads();
                }
                s.base = keep_from;
                s.w = sm.win - s.base;
            }
        } else {
            flush_range(s, s.base, s.count);
            s.in_smem = false;
            s.base = 0;
            s.cap = s.dst_cap;
            s.w = s.dst;
        }
    } else if (want_smem) {
        // re-enter shared memory: bring the live tail back from HBM
        uint32_t n = s.count - keep_from;
        for (uint32_t i = tid; i < n; i += ASM_THREADS) sm.win[i] = s.dst[keep_from + i];
        s.in_smem = true;
        s.base = keep_from;
        s.cap = A.wcap;
        s.w = sm.win - s.base;
    }
    __syncthreads();
}

// ---------------------------------------------------------------- pitch

// estimate_pitch (ctts.c:1899) for two signals of the same length at once (the
// buffer tail `a` and the unit head `b`).
//
// The reference evaluates, for each lag in 55..275, three sequential float sums
// over i < 220: corr += s[i]*s[i+lag], e1 += s[i]^2, e2 += s[i+lag]^2.  The sums
// must keep that order, so the parallel axis is the lag.  Each of 2 x 64 threads
// owns FOUR consecutive lags and walks i in steps of 4 with one 16-byte load of
// x = s[i..i+3] (broadcast) and one of the 4 new y = s[i+lag0+3 .. i+lag0+6]; the
// 7-value sliding window of y (and of y*y, computed once per value) feeds the 16
// (lag, i) pairs of the step: 3.4 instructions per pair instead of 8.5 for the
// one-lag-per-thread loop.  e1 does not depend on the lag: one thread per signal
// computes it concurrently.  The argmax keeps the smallest lag on ties (the
// reference scans lags upward with a strict >).
constexpr int PITCH_Y_WORDS = 576;  // y staging per signal (index + 2 keeps the 16-byte loads aligned)
constexpr int PITCH_X_WORDS = 224;
constexpr int PITCH_SCRATCH_WORDS = 2 * PITCH_Y_WORDS + 2 * PITCH_X_WORDS + 8;

__device__ void estimate_pitch_pair(const Smem& sm, const int16_t* a, const int16_t* b, uint32_t n,
                                    float* pa, float* pb) {
    static_assert(ASM_THREADS >= 192, "needs 128 lag threads + 2 energy threads");
    static_assert((CTTS_PLAN_SAMPLE_RATE / 400) % 4 == 3, "y staging offset assumes min lag = 3 (mod 4)");
    *pa = 0.0f;
    *pb = 0.0f;
    if (n < 200) return;
    const int tid = threadIdx.x;
    const uint32_t lo = CTTS_PLAN_SAMPLE_RATE / 400;
    uint32_t hi = CTTS_PLAN_SAMPLE_RATE / 80;
    if (hi > n / 2) hi = n / 2;
    uint32_t len = CTTS_PLAN_SAMPLE_RATE / 100;
    if (len > n - hi) len = n - hi;
    float* ya = reinterpret_cast<float*>(sm.scratch);
    float* yb = ya + PITCH_Y_WORDS;
    float* xa = yb + PITCH_Y_WORDS;
    float* xb = xa + PITCH_X_WORDS;
    float* e1s = xb + PITCH_X_WORDS;  // [2]
    const uint32_t need = len + hi;   // <= 495
    for (uint32_t i = tid; i < PITCH_Y_WORDS; i += ASM_THREADS) {
        bool in = i >= 2 && i - 2 < need;
        ya[i] = in ? (float)a[i - 2] : 0.0f;
        yb[i] = in ? (float)b[i - 2] : 0.0f;
    }
    for (uint32_t i = tid; i < PITCH_X_WORDS; i += ASM_THREADS) {
        bool in = i < len;
        xa[i] = in ? (float)a[i] : 0.0f;
        xb[i] = in ? (float)b[i] : 0.0f;
    }
    __syncthreads();

    float c[4] = {0.0f, 0.0f, 0.0f, 0.0f}, e2[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    const int sig = (tid >> 6) & 1;
    const uint32_t lag0 = lo + 4u * (uint32_t)(tid & 63);
    if (tid < 128 && lag0 <= hi) {
        const float* x = sig ? xb : xa;
        const float* y = (sig ? yb : ya) + 2 + lag0;  // y[j] = s[lag0 + j]
        float w0 = y[0], w1 = y[1], w2 = y[2];
        float q0 = w0 * w0, q1 = w1 * w1, q2 = w2 * w2;
        const uint32_t len4 = len & ~3u;
        for (uint32_t i = 0; i < len4; i += 4) {
            const float4 xv = *reinterpret_cast<const float4*>(x + i);
            const float4 yn = *reinterpret_cast<const float4*>(y + i + 3);
            const float w3 = yn.x, w4 = yn.y, w5 = yn.z, w6 = yn.w;
            const float q3 = w3 * w3, q4 = w4 * w4, q5 = w5 * w5, q6 = w6 * w6;
            c[0] += xv.x * w0; c[1] += xv.x * w1; c[2] += xv.x * w2; c[3] += xv.x * w3;
            e2[0] += q0; e2[1] += q1; e2[2] += q2; e2[3] += q3;
            c[0] += xv.y * w1; c[1] += xv.y * w2; c[2] += xv.y * w3; c[3] += xv.y * w4;
            e2[0] += q1; e2[1] += q2; e2[2] += q3; e2[3] += q4;
            c[0] += xv.z * w2; c[1] += xv.z * w3; c[2] += xv.z * w4; c[3] += xv.z * w5;
            e2[0] += q2; e2[1] += q3; e2[2] += q4; e2[3] += q5;
            c[0] += xv.w * w3; c[1] += xv.w * w4; c[2] += xv.w * w5; c[3] += xv.w * w6;
            e2[0] += q3; e2[1] += q4; e2[2] += q5; e2[3] += q6;
            w0 = w4; w1 = w5; w2 = w6;
            q0 = q4; q1 = q5; q2 = q6;
        }
        for (uint32_t i = len4; i < len; i++) {
            const float xs = x[i];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const float ys = y[i + k];
                c[k] += xs * ys;
                e2[k] += ys * ys;
            }
        }
    } else if (tid == 128 || tid == 160) {
        const float* x = tid == 160 ? xb : xa;
        float e1 = 0.0f;
        for (uint32_t i = 0; i < len; i++) e1 += x[i] * x[i];
        e1s[tid == 160 ? 1 : 0] = e1;
    }
    __syncthreads();
    unsigned long long key = 0ull;
    if (tid < 128 && lag0 <= hi) {
        const float e1 = e1s[sig];
        float best = 0.0f;
        uint32_t best_lag = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (lag0 + k > hi) break;
            float v = c[k];
            float nrm = sqrtf(e1 * e2[k]);
            if (nrm > 0) v /= nrm;
            if (v > best) {
                best = v;
                best_lag = lag0 + k;
            }
        }
        if (best_lag) key = ((unsigned long long)__float_as_uint(best) << 32) | (0xffffffffu - best_lag);
    }
    // max per signal: warps 0-1 hold signal a, warps 2-3 signal b
    if (tid < 128) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
            key = other > key ? other : key;
        }
        if (lane_id() == 0) sm.red[warp_id()] = key;
    }
    __syncthreads();
    const unsigned long long ka = sm.red[0] > sm.red[1] ? sm.red[0] : sm.red[1];
    const unsigned long long kb = sm.red[2] > sm.red[3] ? sm.red[2] : sm.red[3];
    __syncthreads();
    {
        float v = __uint_as_float((uint32_t)(ka >> 32));
        uint32_t l = 0xffffffffu - (uint32_t)(ka & 0xffffffffu);
        if (ka != 0ull && v > 0.3f && l > 0) *pa = (float)CTTS_PLAN_SAMPLE_RATE / (float)l;
    }
    {
        float v = __uint_as_float((uint32_t)(kb >> 32));
        uint32_t l = 0xffffffffu - (uint32_t)(kb & 0xffffffffu);
        if (kb != 0ull && v > 0.3f && l > 0) *pb = (float)CTTS_PLAN_SAMPLE_RATE / (float)l;
    }
}

// smooth_pitch_boundary + apply_pitch_shift, ctts.c:1946-2024
__device__ void smooth_pitch(const State& s, const Smem& sm, int16_t* us, uint32_t n, uint32_t xf) {
    if (xf == 0 || s.count < 200 || n < 200) return;
    const int tid = threadIdx.x;
    uint32_t reg = xf * 2;
    if (reg > s.count / 2) reg = s.count / 2;
    if (reg > n / 2) reg = n / 2;
    float pp, np;
    estimate_pitch_pair(sm, s.w + (s.count - reg), us, reg, &pp, &np);
    if (!(pp > 0 && np > 0)) return;
    float ratio = np / pp;
    if (!(ratio > 1.15f || ratio < 0.85f)) return;
    float target = (ratio > 1.0f) ? 1.0f + (ratio - 1.0f) * 0.5f : 1.0f - (1.0f - ratio) * 0.5f;
    float shift = target / ratio;
    uint32_t len = xf;
    if (len > n / 4) len = n / 4;
    int16_t* tmp = reinterpret_cast<int16_t*>(sm.scratch);  // len <= ucap/4 samples
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

// match_boundary_energy, ctts.c:1730 (sums of squares are exact integers)
__device__ void match_energy(const State& s, const Smem& sm, int16_t* us, uint32_t n, uint32_t xf) {
    if (xf == 0 || s.count == 0 || n == 0) return;
    const int tid = threadIdx.x;
    uint32_t len = xf;
    if (len > s.count) len = s.count;
    if (len > n) len = n;
    const int16_t* tail = s.w + (s.count - len);
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

// ctts.c:3785-3846: gather -> normalize_rms -> [smooth, match] -> buffer_append_crossfade
__device__ void op_unit(State& s, const Smem& sm, const AsmArgs& A, const ctts_plan_op& op) {
    const int tid = threadIdx.x;
    if (op.a >= A.n_units) { s.err = ERR_BAD_OP; return; }
    const uint32_t n = __ldg(A.unit_cnt + op.a);
    if (n == 0) return;
    if (n > A.ucap) { s.err = ERR_UNIT_TOO_LONG; return; }
    const int16_t* src = A.pool + __ldg(A.unit_off + op.a);
    int16_t* us = sm.ustage;
    const uint32_t xf = op.b;
    const bool boundary = (op.flags & CTTS_UNIT_AFTER_BOUNDARY) != 0;
    const bool join = !boundary && s.count > 0;
    const bool remove_dc = A.prm.remove_dc_offset != 0;
    // samples [0, head) may still be rewritten by smooth_pitch / match_energy
    const uint32_t head = join ? (xf < n ? xf : n) : 0u;

    // gather with 16-byte loads (pool is zero-padded to 8 samples per unit) + sum of squares
    long long ss = 0;
    const uint32_t nvec = (n + 7) >> 3;
    for (uint32_t v = tid; v < nvec; v += ASM_THREADS) {
        int4 q = __ldg(reinterpret_cast<const int4*>(src) + v);
        *(reinterpret_cast<int4*>(us) + v) = q;
        const int16_t* e = reinterpret_cast<const int16_t*>(&q);
#pragma unroll
        for (int k = 0; k < 8; k++) {
            int x = e[k];
            ss += (long long)x * x;
        }
    }
    ss = block_allreduce<ASM_THREADS>(ss, OpAddI64(), reinterpret_cast<long long*>(sm.red));
    // normalize_rms, ctts.c:1709 (double sum of squares == integer sum, exactly); the DC sum of
    // everything past `head` is taken on the way (integer, order-free)
    int dsum = 0;
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
    for (uint32_t v = tid; v < nvec; v += ASM_THREADS) {
        int4 q = *(reinterpret_cast<int4*>(us) + v);
        int16_t* e = reinterpret_cast<int16_t*>(&q);
#pragma unroll
        for (int k = 0; k < 8; k++) {
            int y = e[k];
            if (scale) {
                y = (int)f2s(clamp16f((float)y * g));
                e[k] = (int16_t)y;
            }
            if (v * 8 + k >= head) dsum += y;   // padding past n is zero
        }
        if (scale) *(reinterpret_cast<int4*>(us) + v) = q;
    }
    __syncthreads();

    if (join) {
        smooth_pitch(s, sm, us, n, xf);
        match_energy(s, sm, us, n, xf);
    }

    // buffer_append_crossfade, ctts.c:3279
    int dc = 0;
    if (remove_dc) {
        for (uint32_t i = tid; i < head; i += ASM_THREADS) dsum += us[i];
        long long sum = block_allreduce<ASM_THREADS>((long long)dsum, OpAddI64(), reinterpret_cast<long long*>(sm.red));
        dc = (int)(int16_t)(sum / (long long)n);
    }
    const bool fresh = (s.count == 0) || boundary;
    uint32_t a = 0;
    if (!fresh && xf > 0) {
        a = xf;
        if (a > s.count) a = s.count;
        if (a > n) a = n;
    }
    if ((unsigned long long)s.count - s.base + (n - a) > s.cap) { s.err = ERR_WINDOW_OVERFLOW; return; }
    int16_t* tail = s.w + (s.count - a);   // unit sample i lands at tail[i]

    // prefix that is not a plain copy: the fade-in (ctts.c:3015) or the crossfade mix (ctts.c:3328-3344)
    uint32_t pre = 0;
    if (fresh) {
        pre = A.prm.fade_in_samples < n ? A.prm.fade_in_samples : n;
        if (pre) {
            const float inv = 1.0f / (float)pre;
            for (uint32_t i = tid; i < pre; i += ASM_THREADS) {
                int v = us[i];
                if (remove_dc) v = sub_dc(v, dc);
                tail[i] = f2s((float)v * lut_lerp(A.tab.sine, (float)i * inv));
            }
        }
    } else if (a) {
        pre = a;
        const float inv = 1.0f / (float)a;
        for (uint32_t i = tid; i < a; i += ASM_THREADS) {
            int v = us[i];
            if (remove_dc) v = sub_dc(v, dc);
            float pg, ng;
            crossfade_gains(A.tab, (float)i * inv, &pg, &ng);
            int p = tail[i];
            int mix = (int)((float)p * pg + (float)v * ng);
            tail[i] = (int16_t)max(min(mix, 32767), -32768);
        }
    }
    // the rest: DC removal + copy, 8 samples per thread from the aligned staging buffer
    {
        const uint32_t v0 = pre >> 3;
        for (uint32_t v = v0 + tid; v < nvec; v += ASM_THREADS) {
            int4 q = *(reinterpret_cast<const int4*>(us) + v);
            const int16_t* e = reinterpret_cast<const int16_t*>(&q);
            const uint32_t i0 = v << 3;
            int y[8];
#pragma unroll
            for (int k = 0; k < 8; k++) y[k] = remove_dc ? sub_dc((int)e[k], dc) : (int)e[k];
            if (i0 >= pre && i0 + 8 <= n) {
                int16_t* d = tail + i0;
                if ((reinterpret_cast<uintptr_t>(d) & 3) == 0) {
                    uint32_t* d32 = reinterpret_cast<uint32_t*>(d);
#pragma unroll
                    for (int k = 0; k < 4; k++) d32[k] = (uint32_t)(y[2 * k] & 0xffff) | ((uint32_t)y[2 * k + 1] << 16);
                } else {
                    d[0] = (int16_t)y[0];
                    uint32_t* d32 = reinterpret_cast<uint32_t*>(d + 1);
#pragma unroll
                    for (int k = 0; k < 3; k++)
                        d32[k] = (uint32_t)(y[2 * k + 1] & 0xffff) | ((uint32_t)y[2 * k + 2] << 16);
                    d[7] = (int16_t)y[7];
                }
            } else {
#pragma unroll
                for (int k = 0; k < 8; k++)
                    if (i0 + k >= pre && i0 + k < n) tail[i0 + k] = (int16_t)y[k];
            }
        }
    }
    s.count += n - a;
    __syncthreads();
}

// ---------------------------------------------------------------- word end

// remove_silence_regions, ctts.c:1634, as a bitmask + scan + in-place compaction.
// Returns the new length.  `reg` = w + word_start, len = count - word_start.
__device__ uint32_t trim_region(const Smem& sm, const AsmArgs& A, uint32_t utt, int16_t* reg, uint32_t len) {
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
    else words = A.trim_scratch + (size_t)utt * A.trim_scratch_words;
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
__device__ void op_word_end(State& s, const Smem& sm, const AsmArgs& A, uint32_t utt, const ctts_plan_op& op) {
    const int tid = threadIdx.x;
    if ((op.flags & CTTS_WE_TRIM) && s.count > s.word_start) {
        uint32_t len = s.count - s.word_start;
        if (len > A.prm.min_silence_samples)
            s.count = s.word_start + trim_region(sm, A, utt, s.w + s.word_start, len);
    }
    if (s.count <= s.word_start) return;
    const uint32_t n = s.count - s.word_start;
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

__global__ void __launch_bounds__(ASM_THREADS, 2) assemble_kernel(const AsmArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem sm;
    sm.win = reinterpret_cast<int16_t*>(smem_raw);
    sm.ustage = sm.win + A.wcap;
    sm.scratch = reinterpret_cast<uint32_t*>(sm.ustage + A.ucap);
    sm.hann256 = reinterpret_cast<float*>(sm.scratch + A.scr_words);
    sm.nrm2 = sm.hann256 + PITCH_FRAME;
    sm.red = reinterpret_cast<unsigned long long*>(sm.nrm2 + PITCH_FRAME / 2);

    const int tid = threadIdx.x;
    for (int i = tid; i < PITCH_FRAME; i += ASM_THREADS) sm.hann256[i] = __ldg(A.tab.hann256 + i);
    // norm of a sample covered by two frames: (0 + w[i+128]) + w[i], the reference's accumulation order
    for (int i = tid; i < PITCH_FRAME / 2; i += ASM_THREADS)
        sm.nrm2[i] = __ldg(A.tab.hann256 + i + PITCH_FRAME / 2) + __ldg(A.tab.hann256 + i);

    const UttTask task = A.tasks[blockIdx.x];
    State s;
    s.dst = (task.to_pre ? A.dst_pre : A.dst_final) + task.dst_off;
    s.dst_cap = task.dst_cap;
    s.count = 0;
    s.word_start = 0;
    s.err = 0;
    s.base = 0;
    if ((unsigned long long)task.first_bound + 8ull <= A.wcap) {
        s.in_smem = true;
        s.w = sm.win;
        s.cap = A.wcap;
    } else {
        s.in_smem = false;
        s.w = s.dst;
        s.cap = s.dst_cap;
    }
    __syncthreads();

    for (uint32_t k = task.op_begin; k < task.op_end && !s.err; k++) {
        ctts_plan_op op;
        {
            const int4* p = reinterpret_cast<const int4*>(A.ops + k);
            int4 lo = __ldg(p), hi = __ldg(p + 1);
            *reinterpret_cast<int4*>(&op) = lo;
            *(reinterpret_cast<int4*>(&op) + 1) = hi;
        }
        switch (op.kind) {
            case CTTS_OP_UNIT:
                op_unit(s, sm, A, op);
                break;
            case CTTS_OP_SILENCE: {  // buffer_append_silence, ctts.c:3361
                if ((unsigned long long)s.count - s.base + op.a > s.cap) { s.err = ERR_WINDOW_OVERFLOW; break; }
                int16_t* p = s.w + s.count;
                for (uint32_t i = tid; i < op.a; i += ASM_THREADS) p[i] = 0;
                s.count += op.a;
                __syncthreads();
                break;
            }
            case CTTS_OP_FADE_OUT: {  // apply_fade_out on the buffer tail, ctts.c:3028
                if (s.count > 0 && op.a > 0) {
                    uint32_t f = op.a < s.count ? op.a : s.count;
                    int16_t* p = s.w + (s.count - f);
                    float inv = 1.0f / (float)f;
                    for (uint32_t i = tid; i < f; i += ASM_THREADS)
                        p[i] = f2s((float)p[i] * lut_lerp(A.tab.sine, (float)(f - i) * inv));
                    __syncthreads();
                }
                break;
            }
            case CTTS_OP_WORD_END:
                op_word_end(s, sm, A, task.utt, op);
                break;
            case CTTS_OP_MARK:
                s.word_start = s.count;
                region_switch(s, sm, A, op.a);
                break;
            default:
                s.err = ERR_BAD_OP;
        }
    }
    __syncthreads();
    if (!s.err && s.in_smem) flush_range(s, s.base, s.count);
    if (tid == 0) {
        uint32_t cnt = s.err ? 0u : s.count;
        if (task.to_pre) A.pre_counts[task.utt] = cnt;
        else A.out_counts[task.utt] = cnt;
        A.err[task.utt] = s.err;
    }
}

}  // namespace ctts
