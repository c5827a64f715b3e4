// asm_unit.cuh -- the normalized pool (normalize_rms) and the UNIT op: gather, join, buffer_append_crossfade.
#pragma once
#include "asm_common.cuh"
#include "asm_pitch.cuh"

namespace ctts {

// ---------------------------------------------------------------- unit op

__device__ __forceinline__ int sub_dc(int v, int dc) {
    // clamp(v - dc) to int16 (remove_dc_offset, ctts.c:1577-1581)
    return max(__viaddmin_s32(v, -dc, 32767), -32768);
}

// normalize_rms's per-sample step (ctts.c:1720-1725): (int16)clamp(x * g), truncating
__device__ __forceinline__ uint32_t scale2(uint32_t w, float g) {
    const int y0 = cvt_sat_s16((float)(short)(w & 0xffffu) * g);
    const int y1 = cvt_sat_s16((float)(short)(w >> 16) * g);
    return (uint32_t)(y0 & 0xffff) | ((uint32_t)y1 << 16);
}
__device__ __forceinline__ int4 scale8(const int4& q, float g) {
    int4 r;
    r.x = (int)scale2((uint32_t)q.x, g);
    r.y = (int)scale2((uint32_t)q.y, g);
    r.z = (int)scale2((uint32_t)q.z, g);
    r.w = (int)scale2((uint32_t)q.w, g);
    return r;
}
__device__ __forceinline__ int sum8_s16(const int4& q, int c) {
    c = sum2_s16((uint32_t)q.x, c);
    c = sum2_s16((uint32_t)q.y, c);
    c = sum2_s16((uint32_t)q.z, c);
    return sum2_s16((uint32_t)q.w, c);
}

// remove_dc_offset's per-sample step (ctts.c:1577-1581) on two samples: clamp(v - dc).
// v is first clamped to [lo, hi] = the values whose difference cannot leave the int16 range,
// then a wrapping packed subtract is exact.
struct DcPack { uint32_t lo2, hi2, neg2; };
__device__ __forceinline__ DcPack dc_pack(int dc) {
    const int lo = dc > 0 ? -32768 + dc : -32768;
    const int hi = dc < 0 ? 32767 + dc : 32767;
    DcPack d;
    d.lo2 = (uint32_t)(lo & 0xffff) * 0x10001u;
    d.hi2 = (uint32_t)(hi & 0xffff) * 0x10001u;
    d.neg2 = (uint32_t)((-dc) & 0xffff) * 0x10001u;
    return d;
}
__device__ __forceinline__ uint32_t sub_dc2(uint32_t w, const DcPack& d) {
    return __vadd2(__vmins2(__vmaxs2(w, d.lo2), d.hi2), d.neg2);
}

// normalize_rms (ctts.c:1709, target 3000 at :3684) is the first thing the reference does to
// its private copy of a unit, so its result is a function of the unit and target_rms alone.  It is
// evaluated once per unit when a context first sees a target_rms -- the NORMALIZED POOL, same
// layout as the PCM pool -- instead of once per use (331 000 times per 4096-utterance batch).
// meta[u] = {sum of the normalized samples (lo, hi), (int16)(sum / n) = the DC offset
// remove_dc_offset (ctts.c:1568) finds on the untouched unit, 0}.
__global__ void __launch_bounds__(ASM_THREADS) normalize_pool_kernel(const int16_t* __restrict__ pool, int16_t* __restrict__ out,
                                                                     const uint32_t* __restrict__ unit_off,
                                                                     const uint32_t* __restrict__ unit_cnt, int4* __restrict__ meta,
                                                                     float target_rms) {
    __shared__ long long red[2 * ASM_WARPS + 2];
    const uint32_t u = blockIdx.x;
    const uint32_t n = unit_cnt[u];
    const int tid = threadIdx.x;
    if (n == 0) {
        if (tid == 0) meta[u] = make_int4(0, 0, 0, 0);
        return;
    }
    const int4* srcv = reinterpret_cast<const int4*>(pool + unit_off[u]);
    int4* dstv = reinterpret_cast<int4*>(out + unit_off[u]);
    const uint32_t nvec = (n + 7) >> 3;   // the pool is zero padded to whole vectors; scaling keeps zeros
    // sum of squares: the reference's double sum of integers is the integer sum (< 2^53)
    long long ss = 0;
    for (uint32_t v = tid; v < nvec; v += ASM_THREADS) ss += sumsq8(__ldg(srcv + v));
    // gain = clamp(target / rms), rms = (float)sqrt(sum / n) in double (calculate_rms, ctts.c:1697)
    const unsigned long long gs = block_sum_then<ASM_THREADS>(ss, red, [&](long long tot) {
        float gg = 1.0f;
        uint32_t sc = 0;
        if (target_rms > 0) {
            const float rms = (float)sqrt((double)tot / (double)n);
            if (!(rms < 1.0f)) {
                gg = target_rms / rms;
                if (gg > 3.0f) gg = 3.0f;
                if (gg < 0.1f) gg = 0.1f;
                sc = 1;
            }
        }
        return ((unsigned long long)sc << 32) | __float_as_uint(gg);
    });
    const bool scale = (gs >> 32) != 0;
    const float g = __uint_as_float((uint32_t)gs);
    long long dsum = 0;
    for (uint32_t v = tid; v < nvec; v += ASM_THREADS) {
        int4 q = __ldg(srcv + v);
        if (scale) q = scale8(q, g);
        dstv[v] = q;
        dsum += sum8_s16(q, 0);
    }
    (void)block_sum_then<ASM_THREADS>(dsum, red, [&](long long sum) {
        meta[u] = make_int4((int)(uint32_t)(unsigned long long)sum, (int)(uint32_t)((unsigned long long)sum >> 32),
                            (int)(int16_t)(sum / (long long)n), 0);
        return 0ull;
    });
}

// ctts.c:3785-3846: gather -> normalize_rms -> [smooth, match] -> buffer_append_crossfade.
//
// The gather reads the normalized pool (see normalize_pool_kernel).  The unit's first `hs`
// samples (everything the join may rewrite, and what the pitch analysis reads) are staged in
// `hstage` and joined there; that settles the DC offset (table sum - staged head as gathered +
// staged head as joined), and the rest of the unit then goes from the pool straight to its final
// place in the window, minus the offset, on the window's own 16-byte grid (the pool side is
// re-aligned with a funnel shift): every window sample of the body is written exactly once.
__device__ void op_unit(State& s, const Smem& sm, const AsmArgs& A, const ctts_plan_op& op) {
    const int tid = threadIdx.x;
    if (op.a >= A.n_units) { s.err = ERR_BAD_OP; return; }
    // the plan compiler stored the unit's length and pool offset in the op's unused float fields
    const uint32_t n = __float_as_uint(op.f0);
    if (n == 0) return;
    const int4 meta = __ldg(A.unit_meta + op.a);   // needed only after the join
    // pitch of the unit head over min(2*xf, n/2) samples, if the plan compiler found it in the table
    const uint32_t np_slot = __float_as_uint(op.f2);
    const float np_tab = np_slot ? __ldg(A.unit_pitch + (np_slot - 1)) : 0.0f;
    const int16_t* src = A.pool + __float_as_uint(op.f1);
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
    uint32_t reg_full = 0;      // its value when the buffer is long enough: what the table is for
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
            reg_full = m2;
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

    // unit sample i lands at tail[i]; the body is [hs, n)
    int16_t* tail = s.w + ((int)s.cnt - (int)a);
    const uint32_t phase = (uint32_t)((reinterpret_cast<uintptr_t>(tail + hs) >> 1) & 7u);
    int16_t* grid = tail + hs - phase;                       // 16-byte aligned
    const uint32_t body = n - hsn;                            // may be 0
    const uint32_t gvec = body ? (phase + body + 7) >> 3 : 0; // window vectors that hold body samples
    const uint32_t pv0 = hs >> 3;                             // pool vector of unit sample hs

    // ---- head -> hstage, join there (smooth_pitch_boundary, match_boundary_energy)
    int dc = 0;
    if (hs) {
        int hsum = 0;   // the staged head as gathered (the pool's zero padding adds nothing)
        for (uint32_t v = tid; v < (hs >> 3); v += ASM_THREADS) {
            const int4 q = __ldg(srcv + v);
            *(reinterpret_cast<int4*>(us) + v) = q;
            hsum = sum8_s16(q, hsum);
        }
        __syncthreads();
        smooth_pitch(s, sm, us, n, xf, reg, np_slot != 0 && reg == reg_full, np_tab);
        match_energy(s, sm, us, a);
        // remove_dc_offset (ctts.c:1568) inside buffer_append_crossfade (ctts.c:3279): the sum over the
        // joined unit = table sum - head as gathered + head as joined
        if (remove_dc) {
            int d = -hsum;
            for (uint32_t i = tid; i < hsn; i += ASM_THREADS) d += us[i];
            const long long unit_sum = (long long)(((unsigned long long)(uint32_t)meta.y << 32) | (uint32_t)meta.x);
            dc = (int)block_sum_then<ASM_THREADS>((long long)d, reinterpret_cast<long long*>(sm.red), [&](long long delta) {
                return (unsigned long long)(uint32_t)(int)(int16_t)((unit_sum + delta) / (long long)n);
            });
        }
    } else if (remove_dc) {
        dc = meta.z;   // untouched unit: its DC offset is a table entry
    }

    // ---- body: pool -> minus DC -> window, once
    // (keeping the loads of two vectors in flight, an L1 prefetch before the join and taking tickets
    // one task ahead were all measured slower: the SM is issue bound, other CTAs fill the wait)
    {
        const DcPack dp = dc_pack(dc);
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
            if (dc != 0) {
                q.x = (int)sub_dc2((uint32_t)q.x, dp);
                q.y = (int)sub_dc2((uint32_t)q.y, dp);
                q.z = (int)sub_dc2((uint32_t)q.z, dp);
                q.w = (int)sub_dc2((uint32_t)q.w, dp);
            }
            const int i0 = (int)hs - (int)phase + 8 * (int)j;
            if (i0 >= (int)hs && i0 + 8 <= (int)n) {
                *(reinterpret_cast<int4*>(grid) + j) = q;
            } else {   // the (at most two) partial vectors at the ends of the body
                const int16_t* e = reinterpret_cast<const int16_t*>(&q);
#pragma unroll
                for (int k = 0; k < 8; k++)
                    if (i0 + k >= (int)hs && i0 + k < (int)n) grid[8 * j + k] = e[k];
            }
        }
    }
    if (hs) {
        // staged head: crossfade mix (ctts.c:3328-3344) over [0, a), plain copy over [a, hsn)
        const float inv = a ? 1.0f / (float)a : 0.0f;
        for (uint32_t i = tid; i < a; i += ASM_THREADS) {
            // fast_fade_out / fast_fade_in (ctts.c:76-92) share the table position;
            // xfade4[k] = {fade_out[k], fade_out[k+1], fade_in[k], fade_in[k+1]}: one 16-byte load
            const float x = ((float)i * inv) * (float)(LUT_N - 1);
            const int k = (int)x;
            float pg, ng;
            if (k >= LUT_N - 1 || k < 0) {   // past the ends: the end entry itself
                const float4 e = __ldg(A.tab.xfade4 + (k < 0 ? 0 : LUT_N - 1));
                pg = e.x;
                ng = e.z;
            } else {
                const float4 e = __ldg(A.tab.xfade4 + k);
                const float fr = x - (float)k, om = 1.0f - fr;
                pg = e.x * om + e.y * fr;
                ng = e.z * om + e.w * fr;
            }
            int v = us[i];
            if (remove_dc) v = sub_dc(v, dc);
            const int p = tail[i];
            // clamp((int32)(prev * fo + next * fi)) (ctts.c:3336-3340): truncation then clamp == saturating convert
            tail[i] = (int16_t)cvt_sat_s16((float)p * pg + (float)v * ng);
        }
        for (uint32_t i = a + tid; i < hsn; i += ASM_THREADS) {
            int v = us[i];
            if (remove_dc) v = sub_dc(v, dc);
            tail[i] = (int16_t)v;
        }
    } else if (!join) {
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


}  // namespace ctts
