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
    int16_t* carry;              // 256 samples
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

// estimate_pitch (ctts.c:1899) for two signals of the same length at once:
// threads 0..255 take the lags of `a` (buffer tail), 256..511 those of `b`
// (unit head).  Each lag is one thread's sequential float loop, exactly the
// reference's accumulation order; the argmax keeps the smallest lag on ties
// (the reference scans lags upward with a strict >).
__device__ void estimate_pitch_pair(const Smem& sm, const int16_t* a, const int16_t* b, uint32_t n,
                                    float* pa, float* pb) {
    static_assert(ASM_THREADS == 512, "lag mapping assumes 2 x 256 threads");
    *pa = 0.0f;
    *pb = 0.0f;
    if (n < 200) return;
    const int tid = threadIdx.x;
    uint32_t lo = CTTS_PLAN_SAMPLE_RATE / 400, hi = CTTS_PLAN_SAMPLE_RATE / 80;
    if (hi > n / 2) hi = n / 2;
    uint32_t len = CTTS_PLAN_SAMPLE_RATE / 100;
    if (len > n - hi) len = n - hi;
    float* fa = reinterpret_cast<float*>(sm.scratch);
    float* fb = fa + 512;
    uint32_t need = len + hi;  // <= 495
    for (uint32_t i = tid; i < need; i += ASM_THREADS) {
        fa[i] = (float)a[i];
        fb[i] = (float)b[i];
    }
    __syncthreads();
    const int half = tid >> 8;
    const uint32_t lag = lo + (uint32_t)(tid & 255);
    const float* f = half ? fb : fa;
    unsigned long long key = 0ull;
    if (lag <= hi) {
        float c = 0.0f, e1 = 0.0f, e2 = 0.0f;
        const float* g = f + lag;
#pragma unroll 4
        for (uint32_t i = 0; i < len; i++) {
            float x = f[i], y = g[i];
            c += x * y;
            e1 += x * x;
            e2 += y * y;
        }
        float nrm = sqrtf(e1 * e2);
        if (nrm > 0) c /= nrm;
        if (c > 0.0f) key = ((unsigned long long)__float_as_uint(c) << 32) | (0xffffffffu - lag);
    }
    // max per half
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
        key = other > key ? other : key;
    }
    if (lane_id() == 0) sm.red[warp_id()] = key;
    __syncthreads();
    unsigned long long ka = 0ull, kb = 0ull;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        unsigned long long x = sm.red[k], y = sm.red[8 + k];
        ka = x > ka ? x : ka;
        kb = y > kb ? y : kb;
    }
    __syncthreads();
    {
        float c = __uint_as_float((uint32_t)(ka >> 32));
        uint32_t l = 0xffffffffu - (uint32_t)(ka & 0xffffffffu);
        if (ka != 0ull && c > 0.3f && l > 0) *pa = (float)CTTS_PLAN_SAMPLE_RATE / (float)l;
    }
    {
        float c = __uint_as_float((uint32_t)(kb >> 32));
        uint32_t l = 0xffffffffu - (uint32_t)(kb & 0xffffffffu);
        if (kb != 0ull && c > 0.3f && l > 0) *pb = (float)CTTS_PLAN_SAMPLE_RATE / (float)l;
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
    unsigned long long sp = 0, sn = 0;
    for (uint32_t i = tid; i < len; i += ASM_THREADS) {
        int p = tail[i], q = us[i];
        sp += (unsigned long long)(p * p);
        sn += (unsigned long long)(q * q);
    }
    sp = block_allreduce<ASM_THREADS>(sp, OpAddU64(), sm.red);
    sn = block_allreduce<ASM_THREADS>(sn, OpAddU64(), sm.red);
    float pr = (float)sqrt((double)sp / (double)len);
    float nr = (float)sqrt((double)sn / (double)len);
    if (pr < 1.0f || nr < 1.0f) return;
    float ratio = pr / nr;
    if (ratio > 2.0f) ratio = 2.0f;
    if (ratio < 0.5f) ratio = 0.5f;
    for (uint32_t i = tid; i < len; i += ASM_THREADS) {
        float t = (float)i / (float)len;
        float g = ratio * (1.0f - t) + 1.0f * t;
        us[i] = f2s(clamp16f((float)us[i] * g));
    }
    __syncthreads();
}

// ---------------------------------------------------------------- unit op

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

    // gather with 16-byte loads (pool is zero-padded to 8 samples per unit) + sum of squares
    unsigned long long ss = 0;
    const uint32_t nvec = (n + 7) >> 3;
    for (uint32_t v = tid; v < nvec; v += ASM_THREADS) {
        int4 q = __ldg(reinterpret_cast<const int4*>(src) + v);
        *(reinterpret_cast<int4*>(us) + v) = q;
        const int16_t* e = reinterpret_cast<const int16_t*>(&q);
#pragma unroll
        for (int k = 0; k < 8; k++) {
            int x = e[k];
            ss += (unsigned long long)(x * x);
        }
    }
    ss = block_allreduce<ASM_THREADS>(ss, OpAddU64(), sm.red);
    // normalize_rms, ctts.c:1709 (double sum of squares == integer sum, exactly)
    if (A.prm.target_rms > 0) {
        float rms = (float)sqrt((double)ss / (double)n);
        if (!(rms < 1.0f)) {
            float g = A.prm.target_rms / rms;
            if (g > 3.0f) g = 3.0f;
            if (g < 0.1f) g = 0.1f;
            for (uint32_t v = tid; v < nvec; v += ASM_THREADS) {
                int4 q = *(reinterpret_cast<int4*>(us) + v);
                int16_t* e = reinterpret_cast<int16_t*>(&q);
#pragma unroll
                for (int k = 0; k < 8; k++) e[k] = f2s(clamp16f((float)e[k] * g));
                *(reinterpret_cast<int4*>(us) + v) = q;
            }
        }
    }
    __syncthreads();

    if (!boundary && s.count > 0) {
        smooth_pitch(s, sm, us, n, xf);
        match_energy(s, sm, us, n, xf);
    }

    // buffer_append_crossfade, ctts.c:3279
    int dc = 0;
    if (A.prm.remove_dc_offset) {
        long long sum = 0;
        for (uint32_t i = tid; i < n; i += ASM_THREADS) sum += us[i];
        sum = block_allreduce<ASM_THREADS>(sum, OpAddI64(), reinterpret_cast<long long*>(sm.red));
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
    int16_t* tail = s.w + (s.count - a);
    uint32_t fin = 0;
    float inv = 0.0f;
    if (fresh) {
        fin = A.prm.fade_in_samples < n ? A.prm.fade_in_samples : n;
        if (fin) inv = 1.0f / (float)fin;
    } else if (a) {
        inv = 1.0f / (float)a;
    }
    for (uint32_t i = tid; i < n; i += ASM_THREADS) {
        int v = us[i];
        if (A.prm.remove_dc_offset) {
            v -= dc;
            if (v > 32767) v = 32767;
            if (v < -32768) v = -32768;
        }
        if (fresh) {
            if (i < fin) v = f2s((float)v * lut_lerp(A.tab.sine, (float)i * inv));
        } else if (i < a) {
            float t = (float)i * inv;
            float pg = lut_lerp(A.tab.fade_out, t);
            float ng = lut_lerp(A.tab.fade_in, t);
            int p = tail[i];
            int mix = (int)((float)p * pg + (float)v * ng);
            if (mix > 32767) mix = 32767;
            else if (mix < -32768) mix = -32768;
            v = mix;
        }
        tail[i] = (int16_t)v;
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
// tile by tile, with the 256 originals behind the tile carried in shared memory.
// Reads the reference performs past the end of its heap copy (undefined
// behaviour there, DESIGN.md "Reference UB") yield 0 here.
__device__ void pitch_contour(const Smem& sm, int16_t* x, uint32_t n, float f0, float f1) {
    if (n < 100 || fabsf(f0 - f1) < 0.01f) return;
    if (n < PITCH_FRAME) return;  // no frame fits: every sample keeps its original value
    const int tid = threadIdx.x;
    const uint32_t frames = (n - PITCH_FRAME) / (PITCH_FRAME / 2) + 1;
    const bool degenerate = (n == PITCH_FRAME);  // 1/(n-256) = inf in the reference: NaN indices
    const float inv = 1.0f / (float)(n - PITCH_FRAME);
    const float* hann = sm.hann256;
    int16_t* carry = sm.carry;
    constexpr uint32_t TILE = ASM_THREADS * CONTOUR_KPT;
    for (uint32_t t0 = 0; t0 < n; t0 += TILE) {
        const uint32_t t1 = min(t0 + TILE, n);
        int16_t outv[CONTOUR_KPT];
#pragma unroll
        for (int r = 0; r < CONTOUR_KPT; r++) {
            uint32_t j = t0 + (uint32_t)tid + (uint32_t)r * ASM_THREADS;
            outv[r] = 0;
            if (j >= t1) continue;
            int16_t acc = 0;
            float norm = 0.0f;
            int k1 = (int)(j >> 7);
#pragma unroll
            for (int kk = k1 - 1; kk <= k1; kk++) {
                if (kk < 0 || (uint32_t)kk >= frames) continue;
                uint32_t pos = (uint32_t)kk << 7;
                uint32_t i = j - pos;
                float wv = hann[i];
                float v = 0.0f;
                if (!degenerate) {
                    float t = (float)pos * inv;
                    float st = t * t * (3.0f - 2.0f * t);
                    float pf = f0 + (f1 - f0) * st;
                    float xs = (float)i * pf;
                    uint32_t k = (uint32_t)(unsigned long long)xs;
                    float fr = xs - (float)k;
                    uint32_t q = pos + k;
                    if (k + 1 < PITCH_FRAME) {
                        int16_t s0 = (q < t0) ? carry[q - (t0 - PITCH_FRAME)] : x[q];
                        int16_t s1 = (q + 1 < t0) ? carry[q + 1 - (t0 - PITCH_FRAME)] : x[q + 1];
                        v = (float)s0 * (1.0f - fr) + (float)s1 * fr;
                    } else if (q < n) {
                        int16_t s0 = (q < t0) ? carry[q - (t0 - PITCH_FRAME)] : x[q];
                        v = (float)s0;
                    }
                }
                acc = (int16_t)(acc + f2s(v * wv));
                norm += wv;
            }
            if (norm > 0.01f) outv[r] = f2s(clamp16f((float)acc / norm));
            else outv[r] = x[j];
        }
        // originals the next tile still needs: [t1-256, t1)
        int16_t save = 0;
        if (tid < PITCH_FRAME && t1 >= PITCH_FRAME) save = x[t1 - PITCH_FRAME + tid];
        __syncthreads();
#pragma unroll
        for (int r = 0; r < CONTOUR_KPT; r++) {
            uint32_t j = t0 + (uint32_t)tid + (uint32_t)r * ASM_THREADS;
            if (j < t1) x[j] = outv[r];
        }
        if (tid < PITCH_FRAME) carry[tid] = save;
        __syncthreads();
    }
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
    bool done = false;
    if (op.flags & CTTS_WE_CIRCUMFLEX) {
        uint32_t rise = (uint32_t)(unsigned long long)((float)n * 0.6f);
        if (rise > 100 && n - rise > 100) {
            pitch_contour(sm, x, rise, op.f0, op.f2);
            pitch_contour(sm, x + rise, n - rise, op.f2, op.f1);
            done = true;
        }
    }
    if (!done) pitch_contour(sm, x, n, op.f0, op.f1);
    if (op.flags & CTTS_WE_ENERGY) {
        const float e0 = op.e0, de = op.e1 - op.e0;
        const float den = (float)(n - 1);
        for (uint32_t i = tid; i < n; i += ASM_THREADS) {
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
    sm.carry = reinterpret_cast<int16_t*>(sm.hann256 + PITCH_FRAME);
    sm.red = reinterpret_cast<unsigned long long*>(sm.carry + PITCH_FRAME);

    const int tid = threadIdx.x;
    for (int i = tid; i < PITCH_FRAME; i += ASM_THREADS) sm.hann256[i] = __ldg(A.tab.hann256 + i);

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
