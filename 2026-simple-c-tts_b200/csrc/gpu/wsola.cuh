// wsola.cuh -- time stretching (time_stretch, ctts.c:3490-3617) as two kernels.
//
// 1. wsola_search_kernel: the frame-to-frame dependent chain.  One CTA per
//    stretched utterance walks the frames in order; inside a frame the <= 65
//    coarse candidates (offsets -128..128 step 4) run one per thread, then the
//    <= 6 fine candidates.  Each candidate's correlation is one thread's
//    sequential float loop in the reference's order (groups of 4:
//    ((p0+p1)+p2)+p3, then +=, ctts.c:3411-3413); the coarse candidates share
//    the 4-aligned group sums of squares.  Output: the analysis position of
//    every frame.
// 2. wsola_ola_kernel: embarrassingly parallel gather-form overlap-add.  Each
//    output sample adds its <= 8 windowed frame contributions in frame order
//    into an int16 accumulator that wraps exactly like the reference's `+=`
//    (ctts.c:3577) and a float norm (ctts.c:3578), normalises, and the CTA
//    reports the last non-zero sample for the trailing-zero trim (ctts.c:3611).
//
// This stage is FP32-issue bound (about 13 non-FMA instructions per candidate
// per 4 samples), not HBM bound.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "assemble.cuh"

namespace ctts {

constexpr int WS_FRAME = 512;    // ctts.c:3506
constexpr int WS_HOP = 128;      // analysis hop
constexpr int WS_OVERLAP = 384;  // correlation length
constexpr int WS_SHIFT = 128;    // +-search range
constexpr int WS_THREADS = 128;
constexpr int WS_RANGE = 2 * WS_SHIFT + WS_OVERLAP;  // 640 samples visible to the candidates
constexpr int OLA_THREADS = 256;
constexpr int OLA_SPT = 8;  // samples per thread

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
    uint32_t n_tasks;
    const int16_t* pre;
    const uint32_t* pre_counts;
    int16_t* out;
    uint32_t* out_counts;
    uint32_t* frame_pos;
    uint32_t* n_frames;   // per task
    const float* hann512;
    const uint32_t* ola_block_task;   // per OLA block: task index
    const uint32_t* ola_block_first;  // per OLA block: first output sample
};

__device__ __forceinline__ uint32_t sortable(float v) {
    uint32_t b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

__global__ void __launch_bounds__(WS_THREADS) wsola_search_kernel(const WsolaArgs A) {
    __shared__ __align__(16) float xin[WS_RANGE];    // input[nominal-128 .. nominal+512)
    __shared__ __align__(16) float tgt[WS_OVERLAP];  // previous frame's last 384 samples
    __shared__ float gsq[WS_RANGE / 4];              // 4-aligned group sums of squares of xin
    __shared__ float s_tsq;                          // target sum of squares
    __shared__ unsigned long long s_red[WS_THREADS / 32];
    __shared__ int s_best;

    const StretchTask task = A.tasks[blockIdx.x];
    const int tid = threadIdx.x;
    const uint32_t n = A.pre_counts[task.utt];
    const int16_t* in = A.pre + task.pre_off;
    uint32_t* fpos = A.frame_pos + task.pos_off;

    uint32_t frames = n >= WS_FRAME ? (n - WS_FRAME) / WS_HOP + 1 : 0;
    if (frames > task.max_frames) frames = task.max_frames;
    if (tid == 0) A.n_frames[blockIdx.x] = frames;
    if (frames == 0) return;
    if (tid == 0) fpos[0] = 0;
    uint32_t prev_pos = 0;

    for (uint32_t k = 1; k < frames; k++) {
        const int nominal = (int)(k * WS_HOP);
        // stage the candidates' view and the target as floats
        for (int i = tid; i < WS_RANGE; i += WS_THREADS) {
            int p = nominal - WS_SHIFT + i;
            xin[i] = (p >= 0 && (uint32_t)p < n) ? (float)in[p] : 0.0f;
        }
        for (int i = tid; i < WS_OVERLAP; i += WS_THREADS) tgt[i] = (float)in[prev_pos + (WS_FRAME - WS_OVERLAP) + i];
        __syncthreads();
        for (int g = tid; g < WS_RANGE / 4; g += WS_THREADS) {
            float a0 = xin[4 * g], a1 = xin[4 * g + 1], a2 = xin[4 * g + 2], a3 = xin[4 * g + 3];
            gsq[g] = a0 * a0 + a1 * a1 + a2 * a2 + a3 * a3;
        }
        if (tid == WS_THREADS - 1) {
            float sb = 0.0f;
            for (int m = 0; m < WS_OVERLAP / 4; m++) {
                float b0 = tgt[4 * m], b1 = tgt[4 * m + 1], b2 = tgt[4 * m + 2], b3 = tgt[4 * m + 3];
                sb += b0 * b0 + b1 * b1 + b2 * b2 + b3 * b3;
            }
            s_tsq = sb;
        }
        __syncthreads();

        // coarse: candidate c has offset -128 + 4c
        unsigned long long key = 0ull;
        if (tid <= 2 * WS_SHIFT / 4) {
            int off = -WS_SHIFT + 4 * tid;
            int cpos = nominal + off;
            if (cpos >= 0 && (uint32_t)cpos + WS_FRAME <= n) {
                float sp = 0.0f, sa = 0.0f;
                const float4* xa = reinterpret_cast<const float4*>(xin) + tid;
                const float4* xb = reinterpret_cast<const float4*>(tgt);
                const float* gq = gsq + tid;
#pragma unroll 4
                for (int m = 0; m < WS_OVERLAP / 4; m++) {
                    float4 a = xa[m], b = xb[m];
                    sp += a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
                    sa += gq[m];
                }
                float den = sqrtf(sa * s_tsq);
                float c = den < 1.0f ? 0.0f : sp / den;
                key = ((unsigned long long)sortable(c) << 32) | (unsigned long long)(0xffffu - (uint32_t)tid);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
            key = other > key ? other : key;
        }
        if (lane_id() == 0) s_red[warp_id()] = key;
        __syncthreads();
        unsigned long long best = 0ull;
#pragma unroll
        for (int w = 0; w < WS_THREADS / 32; w++) best = s_red[w] > best ? s_red[w] : best;
        int best_off = 0;
        uint32_t best_corr_key = sortable(-2.0f);
        if (best != 0ull) {
            best_off = -WS_SHIFT + 4 * (int)(0xffffu - (uint32_t)(best & 0xffffu));
            best_corr_key = (uint32_t)(best >> 32);
        }
        __syncthreads();

        // fine: best_off-3 .. best_off+3 without best_off, ascending, strict >
        int lo = best_off - 3, hi = best_off + 3;
        if (lo < -WS_SHIFT) lo = -WS_SHIFT;
        if (hi > WS_SHIFT) hi = WS_SHIFT;
        unsigned long long fkey = 0ull;
        if (tid <= hi - lo) {
            int off = lo + tid;
            int cpos = nominal + off;
            if (off != best_off && cpos >= 0 && (uint32_t)cpos + WS_FRAME <= n) {
                float sp = 0.0f, sa = 0.0f;
                const float* xa = xin + (off + WS_SHIFT);
#pragma unroll 2
                for (int m = 0; m < WS_OVERLAP / 4; m++) {
                    float a0 = xa[4 * m], a1 = xa[4 * m + 1], a2 = xa[4 * m + 2], a3 = xa[4 * m + 3];
                    float b0 = tgt[4 * m], b1 = tgt[4 * m + 1], b2 = tgt[4 * m + 2], b3 = tgt[4 * m + 3];
                    sp += a0 * b0 + a1 * b1 + a2 * b2 + a3 * b3;
                    sa += a0 * a0 + a1 * a1 + a2 * a2 + a3 * a3;
                }
                float den = sqrtf(sa * s_tsq);
                float c = den < 1.0f ? 0.0f : sp / den;
                uint32_t ck = sortable(c);
                if (ck > best_corr_key) fkey = ((unsigned long long)ck << 32) | (unsigned long long)(0xffu - (uint32_t)tid);
            }
        }
        if (tid < 32) {
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) {
                unsigned long long other = __shfl_xor_sync(0xffffffffu, fkey, o);
                fkey = other > fkey ? other : fkey;
            }
            if (tid == 0) {
                int off = best_off;
                if (fkey != 0ull) off = lo + (int)(0xffu - (uint32_t)(fkey & 0xffu));
                s_best = off;
            }
        }
        __syncthreads();
        int off = s_best;
        uint32_t pos = (uint32_t)(nominal + off);
        if (pos + WS_FRAME > n) pos = n - WS_FRAME;
        if (tid == 0) fpos[k] = pos;
        prev_pos = pos;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(OLA_THREADS) wsola_ola_kernel(const WsolaArgs A) {
    __shared__ int s_last[OLA_THREADS / 32];
    __shared__ float win[WS_FRAME];
    const int tid = threadIdx.x;
    for (int i = tid; i < WS_FRAME; i += OLA_THREADS) win[i] = __ldg(A.hann512 + i);
    const uint32_t ti = A.ola_block_task[blockIdx.x];
    const StretchTask task = A.tasks[ti];
    const uint32_t frames = A.n_frames[ti];
    __syncthreads();
    if (frames == 0) return;
    const uint32_t hop = task.hop;
    const uint32_t used = (frames - 1) * hop + WS_FRAME;
    const int16_t* in = A.pre + task.pre_off;
    const uint32_t* fpos = A.frame_pos + task.pos_off;
    int16_t* out = A.out + task.out_off;
    const uint32_t first = A.ola_block_first[blockIdx.x];
    int last_nz = -1;
#pragma unroll
    for (int r = 0; r < OLA_SPT; r++) {
        uint32_t j = first + (uint32_t)r * OLA_THREADS + (uint32_t)tid;
        if (j >= used || j >= task.out_cap) continue;
        uint32_t k_hi = j / hop;
        if (k_hi > frames - 1) k_hi = frames - 1;
        uint32_t k_lo = j < WS_FRAME ? 0u : (j - WS_FRAME) / hop + 1u;
        int16_t acc = 0;
        float norm = 0.0f;
        for (uint32_t k = k_lo; k <= k_hi; k++) {
            uint32_t i = j - k * hop;
            float wv = win[i];
            float v = (float)in[__ldg(fpos + k) + i] * wv;
            acc = (int16_t)(acc + f2s(v));
            norm += wv;
        }
        int16_t y = acc;
        if (norm > 0.01f) y = f2s(clamp16f((float)acc / norm));
        out[j] = y;
        if (y != 0) last_nz = (int)j;
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
