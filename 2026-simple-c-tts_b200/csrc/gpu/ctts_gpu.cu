// ctts_gpu.cu -- C-ABI (include/ctts_gpu.h) over the sm_100a kernels.
//
// Host side only does plumbing: parse voice.db, re-pack and upload the PCM
// pool and tables, turn a CSR plan into per-utterance tasks and an output
// layout, launch kernels on one stream.  There is no CPU implementation of any
// sample operation in this library.
#include "ctts_gpu.h"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <vector>

#include <cuda_runtime.h>

#include "assemble.cuh"
#include "wsola.cuh"

extern "C" void ctts_host_tables(float* fade_out, float* fade_in, float* sine, float* hann256,
                                 float* hann512);

namespace {

struct DbHeader {  // ctts.h:84-98
    uint32_t magic, version, unit_count, sample_rate, bits_per_sample, index_offset, strings_offset,
        audio_offset, total_samples, max_unit_chars, hash_table_size, hash_table_offset;
    uint8_t reserved[16];
};
struct DbEntry {  // ctts.h:101-111
    uint32_t hash, string_offset;
    uint16_t string_len, char_count;
    uint32_t audio_offset, sample_count, flags, next_hash, reserved;
};
static_assert(sizeof(DbHeader) == 64 && sizeof(DbEntry) == 32, "voice.db records");
static_assert(sizeof(ctts_plan_op) == 32, "plan op");

inline uint64_t up8(uint64_t v) { return (v + 7) & ~7ull; }

}  // namespace

struct ctts_gpu_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    int16_t* d_pool = nullptr;
    uint32_t* d_unit_off = nullptr;
    uint32_t* d_unit_cnt = nullptr;
    float* d_tables = nullptr;  // fade_out, fade_in, sine (1024 each), hann256, hann512, xfade4 (4096)
    uint32_t n_units = 0;
    uint32_t max_unit = 0;
    std::vector<uint32_t> unit_cnt;
    int sm_count = 0;
    int smem_per_sm = 0;
    int smem_optin = 0;
    int16_t* d_batch_out = nullptr;  // reused by ctts_gpu_synth_batch
    uint64_t batch_out_cap = 0;
    char err[512] = {0};
};

struct ctts_gpu_plan {
    ctts_gpu_ctx* ctx = nullptr;
    uint32_t n_utts = 0;
    ctts_assembly_params prm{};
    std::vector<uint64_t> offsets;  // n_utts + 1
    std::vector<uint64_t> bounds;
    ctts_plan_op* d_ops = nullptr;
    ctts::RegionTask* d_tasks = nullptr;
    unsigned long long* d_chain = nullptr;
    uint32_t* d_ticket = nullptr;
    uint32_t n_tasks = 0;
    uint32_t n_global_tasks = 0;
    uint32_t epoch = 0;
    uint32_t grid = 0;
    ctts::StretchTask* d_stasks = nullptr;
    uint32_t* d_counts = nullptr;
    uint32_t* d_pre_counts = nullptr;
    uint32_t* d_err = nullptr;
    uint32_t* d_trim = nullptr;
    uint32_t trim_words = 0;
    int16_t* d_pre = nullptr;
    uint32_t* d_frame_pos = nullptr;
    uint32_t* d_n_frames = nullptr;
    uint32_t* d_ola_task = nullptr;
    uint32_t* d_ola_first = nullptr;
    uint32_t n_stretch = 0;
    uint32_t n_ola_blocks = 0;
    int16_t* d_out_owned = nullptr;
    int16_t* d_out_last = nullptr;
    std::vector<uint64_t> pre_off;   // per utterance (stretch only), else ~0
    std::vector<uint64_t> pre_cap;
    uint32_t wcap = 0, hcap = 0, scr_words = 0, smem_bytes = 0;
    ctts_gpu_run_info info{};
};

namespace {

int fail(ctts_gpu_ctx* c, int code, const char* fmt, ...) {
    if (c) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(c->err, sizeof c->err, fmt, ap);
        va_end(ap);
    }
    return code;
}

#define CU(ctx, call)                                                                          \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess)                                                                 \
            return fail((ctx), CTTS_GPU_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                                   \
    } while (0)

template <typename T>
int upload(ctts_gpu_ctx* ctx, T** dptr, const std::vector<T>& h) {
    size_t bytes = std::max<size_t>(h.size(), 1) * sizeof(T);
    CU(ctx, cudaMalloc(reinterpret_cast<void**>(dptr), bytes));
    if (!h.empty())
        CU(ctx, cudaMemcpyAsync(*dptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}

// speed handling of ctts_synthesize / time_stretch (ctts.c:3907, :3493-3503):
// returns true when the utterance goes through WSOLA and sets the synthesis hop
bool needs_stretch(float speed, uint32_t* hop) {
    if (speed == 1.0f) return false;
    if (speed < 0.5f) speed = 0.5f;
    if (speed > 2.0f) speed = 2.0f;
    if (fabsf(speed - 1.0f) < 0.01f) return false;  // plain copy
    size_t h = (size_t)((float)(size_t)128 / speed);
    if (h < 1) h = 1;
    *hop = (uint32_t)h;
    return true;
}

uint64_t stretch_bound(uint64_t pre, uint32_t hop) {
    uint64_t frames = pre > 512 ? (pre - 512) / 128 + 1 : 1;
    return frames * hop + 512;
}

int compute_bounds(const ctts_gpu_ctx* ctx, const ctts_batch_plan* plan, std::vector<uint64_t>* pre,
                   std::vector<uint64_t>* out) {
    pre->assign(plan->n_utts, 0);
    out->assign(plan->n_utts, 0);
    for (uint32_t u = 0; u < plan->n_utts; u++) {
        uint32_t b = plan->utt_op_begin[u], e = plan->utt_op_begin[u + 1];
        if (b > e || e > plan->n_ops) return CTTS_GPU_ERR_INVALID_ARG;
        uint64_t total = 0;
        for (uint32_t k = b; k < e; k++) {
            const ctts_plan_op& op = plan->ops[k];
            if (op.kind == CTTS_OP_UNIT) {
                if (op.a >= ctx->n_units) return CTTS_GPU_ERR_INVALID_ARG;
                total += ctx->unit_cnt[op.a];
            } else if (op.kind == CTTS_OP_SILENCE) {
                total += op.a;
            } else if (op.kind < CTTS_OP_UNIT || op.kind > CTTS_OP_MARK) {
                return CTTS_GPU_ERR_INVALID_ARG;
            }
        }
        (*pre)[u] = total;
        uint32_t hop = 0;
        (*out)[u] = needs_stretch(plan->speed[u], &hop) ? stretch_bound(total, hop) : total;
    }
    return 0;
}

}  // namespace

extern "C" {

int ctts_gpu_init(ctts_gpu_ctx** out, const void* voice_db, size_t db_size, int device_ordinal) {
    if (!out || !voice_db || db_size < sizeof(DbHeader)) return CTTS_GPU_ERR_INVALID_ARG;
    *out = nullptr;
    DbHeader h;
    memcpy(&h, voice_db, sizeof h);
    if (h.magic != 0x53545443u) return CTTS_GPU_ERR_INVALID_FORMAT;
    if (h.version != 1u) return CTTS_GPU_ERR_VERSION;
    if ((uint64_t)h.index_offset + (uint64_t)h.unit_count * sizeof(DbEntry) > db_size ||
        (uint64_t)h.audio_offset + 2ull * h.total_samples > db_size)
        return CTTS_GPU_ERR_INVALID_FORMAT;

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device_ordinal < 0 || device_ordinal >= ndev)
        return CTTS_GPU_ERR_CUDA;  // no CPU fallback by design

    ctts_gpu_ctx* ctx = new ctts_gpu_ctx();
    ctx->device = device_ordinal;
    auto bail = [&](int code) {
        ctts_gpu_free(ctx);
        return code;
    };
#define CUI(call)                                                            \
    do {                                                                     \
        cudaError_t e_ = (call);                                             \
        if (e_ != cudaSuccess) {                                             \
            fprintf(stderr, "ctts_gpu_init: %s: %s\n", #call, cudaGetErrorString(e_)); \
            return bail(CTTS_GPU_ERR_CUDA);                                  \
        }                                                                    \
    } while (0)
    CUI(cudaSetDevice(device_ordinal));
    CUI(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
    ctx->stream = ctx->own_stream;
    CUI(cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device_ordinal));
    CUI(cudaDeviceGetAttribute(&ctx->smem_per_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, device_ordinal));
    CUI(cudaDeviceGetAttribute(&ctx->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device_ordinal));

    // re-pack: every unit starts on a 16-byte boundary, zero padded (int16x8 loads)
    const uint8_t* base = static_cast<const uint8_t*>(voice_db);
    const uint8_t* pcm = base + h.audio_offset;  // may be 2-byte misaligned
    ctx->n_units = h.unit_count;
    ctx->unit_cnt.resize(h.unit_count);
    std::vector<uint32_t> unit_off(h.unit_count);
    uint64_t packed = 0;
    for (uint32_t u = 0; u < h.unit_count; u++) {
        DbEntry e;
        memcpy(&e, base + h.index_offset + (size_t)u * sizeof(DbEntry), sizeof e);
        if ((uint64_t)e.audio_offset + e.sample_count > h.total_samples) return bail(CTTS_GPU_ERR_INVALID_FORMAT);
        ctx->unit_cnt[u] = e.sample_count;
        unit_off[u] = (uint32_t)packed;
        packed += up8(e.sample_count);
        ctx->max_unit = std::max(ctx->max_unit, e.sample_count);
        if (packed > 0xffffffffull) return bail(CTTS_GPU_ERR_INVALID_FORMAT);
    }
    std::vector<int16_t> pool(std::max<uint64_t>(packed, 8), 0);
    for (uint32_t u = 0; u < h.unit_count; u++) {
        DbEntry e;
        memcpy(&e, base + h.index_offset + (size_t)u * sizeof(DbEntry), sizeof e);
        memcpy(pool.data() + unit_off[u], pcm + 2ull * e.audio_offset, 2ull * e.sample_count);
    }
    CUI(cudaMalloc(reinterpret_cast<void**>(&ctx->d_pool), pool.size() * sizeof(int16_t)));
    CUI(cudaMemcpy(ctx->d_pool, pool.data(), pool.size() * sizeof(int16_t), cudaMemcpyHostToDevice));
    CUI(cudaMalloc(reinterpret_cast<void**>(&ctx->d_unit_off), std::max<size_t>(h.unit_count, 1) * 4));
    CUI(cudaMalloc(reinterpret_cast<void**>(&ctx->d_unit_cnt), std::max<size_t>(h.unit_count, 1) * 4));
    if (h.unit_count) {
        CUI(cudaMemcpy(ctx->d_unit_off, unit_off.data(), h.unit_count * 4ull, cudaMemcpyHostToDevice));
        CUI(cudaMemcpy(ctx->d_unit_cnt, ctx->unit_cnt.data(), h.unit_count * 4ull, cudaMemcpyHostToDevice));
    }
    std::vector<float> tab(3 * 1024 + 256 + 512 + 4 * 1024);
    ctts_host_tables(tab.data(), tab.data() + 1024, tab.data() + 2048, tab.data() + 3072, tab.data() + 3328);
    for (int k = 0; k < 1024; k++) {   // interleaved crossfade table: one 16-byte load per sample
        const int k1 = k + 1 < 1024 ? k + 1 : 1023;
        float* e = tab.data() + 3840 + 4 * k;
        e[0] = tab[k];
        e[1] = tab[k1];
        e[2] = tab[1024 + k];
        e[3] = tab[1024 + k1];
    }
    CUI(cudaMalloc(reinterpret_cast<void**>(&ctx->d_tables), tab.size() * sizeof(float)));
    CUI(cudaMemcpy(ctx->d_tables, tab.data(), tab.size() * sizeof(float), cudaMemcpyHostToDevice));
#undef CUI
    *out = ctx;
    return CTTS_GPU_OK;
}

void ctts_gpu_free(ctts_gpu_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->own_stream) cudaStreamSynchronize(ctx->own_stream);
    cudaFree(ctx->d_pool);
    cudaFree(ctx->d_unit_off);
    cudaFree(ctx->d_unit_cnt);
    cudaFree(ctx->d_tables);
    cudaFree(ctx->d_batch_out);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

int ctts_gpu_set_stream(ctts_gpu_ctx* ctx, void* cuda_stream) {
    if (!ctx) return CTTS_GPU_ERR_INVALID_ARG;
    ctx->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->own_stream;
    return CTTS_GPU_OK;
}

const char* ctts_gpu_last_error(const ctts_gpu_ctx* ctx) { return ctx ? ctx->err : "no context"; }

int ctts_gpu_plan_bounds(const ctts_gpu_ctx* ctx, const ctts_batch_plan* plan, uint64_t* out_bound) {
    if (!ctx || !plan || !out_bound) return CTTS_GPU_ERR_INVALID_ARG;
    std::vector<uint64_t> pre, out;
    int rc = compute_bounds(ctx, plan, &pre, &out);
    if (rc) return rc;
    std::copy(out.begin(), out.end(), out_bound);
    return CTTS_GPU_OK;
}

void ctts_gpu_plan_destroy(ctts_gpu_plan* p) {
    if (!p) return;
    if (p->ctx) {
        cudaSetDevice(p->ctx->device);
        cudaStreamSynchronize(p->ctx->stream);
    }
    cudaFree(p->d_ops);
    cudaFree(p->d_tasks);
    cudaFree(p->d_chain);
    cudaFree(p->d_ticket);
    cudaFree(p->d_stasks);
    cudaFree(p->d_counts);
    cudaFree(p->d_pre_counts);
    cudaFree(p->d_err);
    cudaFree(p->d_trim);
    cudaFree(p->d_pre);
    cudaFree(p->d_frame_pos);
    cudaFree(p->d_n_frames);
    cudaFree(p->d_ola_task);
    cudaFree(p->d_ola_first);
    cudaFree(p->d_out_owned);
    delete p;
}

int ctts_gpu_plan_create(ctts_gpu_ctx* ctx, const ctts_batch_plan* plan, const ctts_assembly_params* params,
                         const uint64_t* out_offsets, ctts_gpu_plan** out) {
    if (!ctx || !plan || !params || !out) return CTTS_GPU_ERR_INVALID_ARG;
    if (plan->n_utts && (!plan->utt_op_begin || !plan->speed)) return CTTS_GPU_ERR_INVALID_ARG;
    if (plan->n_ops && !plan->ops) return CTTS_GPU_ERR_INVALID_ARG;
    if (params->min_silence_samples < 10)
        // below 10 the reference's keep = max(min/4, 10) overruns the silent run
        return fail(ctx, CTTS_GPU_ERR_INVALID_ARG, "min_silence_samples < 10 is not memory-safe in the reference");
    *out = nullptr;
    CU(ctx, cudaSetDevice(ctx->device));

    std::vector<uint64_t> pre, bound;
    int rc = compute_bounds(ctx, plan, &pre, &bound);
    if (rc) return fail(ctx, rc, "invalid plan");

    ctts_gpu_plan* p = new ctts_gpu_plan();
    p->ctx = ctx;
    p->n_utts = plan->n_utts;
    p->prm = *params;
    p->bounds = bound;
    const uint32_t n = plan->n_utts;

    // output layout
    p->offsets.resize((size_t)n + 1);
    if (out_offsets) {
        for (uint32_t u = 0; u <= n; u++) p->offsets[u] = out_offsets[u];
        for (uint32_t u = 0; u < n; u++) {
            if ((out_offsets[u] & 7) || out_offsets[u + 1] < out_offsets[u] ||
                out_offsets[u + 1] - out_offsets[u] < bound[u] || out_offsets[u + 1] - out_offsets[u] > 0xffffffffull) {
                ctts_gpu_plan_destroy(p);
                return fail(ctx, CTTS_GPU_ERR_BOUNDS, "output slot %u is misaligned or smaller than its bound %llu", u,
                            (unsigned long long)bound[u]);
            }
        }
    } else {
        uint64_t o = 0;
        for (uint32_t u = 0; u < n; u++) {
            p->offsets[u] = o;
            o += up8(bound[u]) + 8;
        }
        p->offsets[n] = o;
    }

    // ---- plan compile step 1: private copy of the ops; fade-outs that provably act on zeros
    // (or on an empty buffer) become no-ops, so that a pause-only region never has to reach
    // back into its predecessor's samples.  Trailing zeros: appended silence stays zero under
    // apply_fade_out (0 * g == 0); a unit, or a WORD_END over a region that holds audio
    // (trimming / the contour may move samples into the tail), resets the count.
    std::vector<ctts_plan_op> ops(plan->ops, plan->ops + plan->n_ops);
    uint32_t xf_max = 0;
    uint64_t gather = 0;
    bool bad_factor = false;
    for (uint32_t u = 0; u < n; u++) {
        uint64_t tz = 0, count_ub = 0;
        bool audio = false;
        for (uint32_t k = plan->utt_op_begin[u]; k < plan->utt_op_begin[u + 1]; k++) {
            ctts_plan_op& op = ops[k];
            switch (op.kind) {
                case CTTS_OP_UNIT:
                    count_ub += ctx->unit_cnt[op.a];
                    gather += ctx->unit_cnt[op.a];
                    if (ctx->unit_cnt[op.a]) { tz = 0; audio = true; }
                    xf_max = std::max(xf_max, op.b);
                    break;
                case CTTS_OP_SILENCE:
                    tz += op.a;
                    count_ub += op.a;
                    break;
                case CTTS_OP_FADE_OUT:
                    if (count_ub == 0 || tz >= op.a) op.kind = ctts::OP_NOP;
                    break;
                case CTTS_OP_WORD_END:
                    if (audio) tz = 0;
                    if (op.flags & CTTS_WE_INTON) {
                        // the contour kernel stages CONTOUR_AHEAD samples past a tile: factors come from
                        // clamp_pitch(1 +- max_pitch_change) (ctts.c:2589), 0.9 .. 1.1 as shipped
                        const float lo = std::min(op.f0, std::min(op.f1, op.f2)), hi = std::max(op.f0, std::max(op.f1, op.f2));
                        if (!(lo >= 0.0f) || !(hi <= 2.05f)) bad_factor = true;
                    }
                    break;
                case CTTS_OP_MARK:
                    audio = false;
                    break;
            }
        }
    }

    if (bad_factor) {
        ctts_gpu_plan_destroy(p);
        return fail(ctx, CTTS_GPU_ERR_INVALID_ARG, "WORD_END pitch factors must lie in [0, 2.05]");
    }

    // ---- shared-memory geometry
    const uint32_t max_unit = ctx->max_unit;
    const uint32_t hcap = (uint32_t)up8(std::max<uint32_t>(std::min(xf_max, max_unit), 496)) + 8;
    if (hcap > 2 * ctts::SCR_WORDS) {
        ctts_gpu_plan_destroy(p);
        return fail(ctx, CTTS_GPU_ERR_INVALID_ARG, "crossfade of %u samples exceeds the staging capacity", xf_max);
    }
    auto smem_for = [&](uint32_t wcap) { return ctts::SMEM_HSTAGE + hcap * 2 + (wcap + 16) * 2; };
    // regions (the samples between two word marks) and their upper bounds
    struct Region { uint32_t op_begin, op_end; uint64_t bound; uint32_t units; };
    std::vector<std::vector<Region>> regions(n);
    uint64_t region_max = 0;
    for (uint32_t u = 0; u < n; u++) {
        Region r{plan->utt_op_begin[u], plan->utt_op_begin[u], 0, 0};
        for (uint32_t k = plan->utt_op_begin[u]; k < plan->utt_op_begin[u + 1]; k++) {
            const ctts_plan_op& op = ops[k];
            if (op.kind == CTTS_OP_UNIT) { r.bound += ctx->unit_cnt[op.a]; r.units++; }
            else if (op.kind == CTTS_OP_SILENCE) r.bound += op.a;
            r.op_end = k + 1;
            if (op.kind == CTTS_OP_MARK) {
                regions[u].push_back(r);
                region_max = std::max(region_max, r.bound);
                r = Region{k + 1, k + 1, 0, 0};
            }
        }
        if (r.op_end > r.op_begin) {
            regions[u].push_back(r);
            region_max = std::max(region_max, r.bound);
        }
    }
    // window: as large as the target occupancy allows, no larger than the largest region needs
    int want_ctas = 3;
    if (const char* e = getenv("CTTS_GPU_CTAS_PER_SM")) want_ctas = std::max(1, std::min(8, atoi(e)));
    const uint32_t budget = std::min<uint32_t>((uint32_t)ctx->smem_optin, (uint32_t)(ctx->smem_per_sm / want_ctas - 1024));
    uint32_t wcap = (uint32_t)up8(std::min<uint64_t>(region_max + 16, 1u << 20));
    if (const char* e = getenv("CTTS_GPU_WINDOW")) wcap = (uint32_t)up8(std::max(256, atoi(e)));   // tests: force the HBM path
    while (wcap > 1024 && smem_for(wcap) > budget) wcap -= 256;
    wcap &= ~7u;
    if (smem_for(wcap) > (uint32_t)ctx->smem_optin) {
        ctts_gpu_plan_destroy(p);
        return fail(ctx, CTTS_GPU_ERR_INVALID_ARG, "crossfade / unit sizes (%u, %u samples) do not fit shared memory", xf_max, max_unit);
    }
    p->wcap = wcap;
    p->hcap = hcap;
    p->scr_words = ctts::SCR_WORDS;
    p->smem_bytes = smem_for(wcap);

    // ---- plan compile step 2: region tasks.  A region with no unit (a pause) or a tiny one is
    // appended to the task before it while the sum still fits the window.
    struct HostTask { uint32_t utt, op_begin, op_end; uint64_t bound; uint32_t index_in_utt; uint32_t region_max; };
    std::vector<std::vector<HostTask>> utt_tasks(n);
    uint32_t max_tasks_per_utt = 0;
    uint64_t n_tasks = 0;
    for (uint32_t u = 0; u < n; u++) {
        auto& T = utt_tasks[u];
        for (const Region& r : regions[u]) {
            const bool tiny = r.units == 0 || r.bound <= 2048;
            if (!T.empty() && tiny && T.back().bound + r.bound <= wcap) {
                T.back().op_end = r.op_end;
                T.back().bound += r.bound;
                T.back().region_max = (uint32_t)std::max<uint64_t>(T.back().region_max, r.bound);
            } else {
                T.push_back(HostTask{u, r.op_begin, r.op_end, r.bound, (uint32_t)T.size(),
                                     (uint32_t)std::min<uint64_t>(r.bound, 0xffffffffull)});
            }
        }
        max_tasks_per_utt = std::max<uint32_t>(max_tasks_per_utt, (uint32_t)T.size());
        n_tasks += T.size();
    }
    if (n_tasks > 0x7fffffffull) {
        ctts_gpu_plan_destroy(p);
        return fail(ctx, CTTS_GPU_ERR_INVALID_ARG, "too many region tasks");
    }

    // slots: pre-stretch buffers and WSOLA tasks (utterances longest first)
    std::vector<ctts::StretchTask> stasks;
    std::vector<uint32_t> ola_task, ola_first;
    p->pre_off.assign(n, ~0ull);
    p->pre_cap.assign(n, 0);
    uint64_t pre_total = 0, pos_total = 0;
    std::vector<uint32_t> order(n);
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return pre[a] > pre[b]; });
    std::vector<uint64_t> slot_off(n);
    std::vector<uint32_t> slot_cap(n), to_pre(n, 0);
    for (uint32_t i = 0; i < n; i++) {
        uint32_t u = order[i];
        uint32_t hop = 0;
        if (needs_stretch(plan->speed[u], &hop)) {
            if (pre[u] + 16 > 0xffffffffull) {
                ctts_gpu_plan_destroy(p);
                return fail(ctx, CTTS_GPU_ERR_INVALID_ARG, "utterance %u too long", u);
            }
            to_pre[u] = 1;
            slot_off[u] = pre_total;
            slot_cap[u] = (uint32_t)(up8(pre[u]) + 8);
            p->pre_off[u] = pre_total;
            p->pre_cap[u] = slot_cap[u];
            ctts::StretchTask st;
            st.utt = u;
            st.hop = hop;
            st.pre_off = pre_total;
            st.out_off = p->offsets[u];
            st.out_cap = (uint32_t)(p->offsets[u + 1] - p->offsets[u]);
            st.pos_off = (uint32_t)pos_total;
            st.max_frames = (uint32_t)(pre[u] > 512 ? (pre[u] - 512) / 128 + 1 : 1);
            uint64_t used_max = (uint64_t)st.max_frames * hop + 512;
            uint32_t per_block = ctts::OLA_THREADS * ctts::OLA_SPT;
            for (uint64_t f = 0; f < used_max; f += per_block) {
                ola_task.push_back((uint32_t)stasks.size());
                ola_first.push_back((uint32_t)f);
            }
            stasks.push_back(st);
            pre_total += slot_cap[u];
            pos_total += st.max_frames;
            if (pos_total > 0xffffffffull) {
                ctts_gpu_plan_destroy(p);
                return fail(ctx, CTTS_GPU_ERR_INVALID_ARG, "too many WSOLA frames in one batch");
            }
        } else {
            slot_off[u] = p->offsets[u];
            slot_cap[u] = (uint32_t)(p->offsets[u + 1] - p->offsets[u]);
        }
    }
    p->n_stretch = (uint32_t)stasks.size();
    p->n_ola_blocks = (uint32_t)ola_task.size();

    // ticket order: region-major (task k of every utterance before task k+1 of any), so that a
    // task's predecessor has normally finished long before the task starts
    std::vector<ctts::RegionTask> tasks;
    tasks.reserve(n_tasks);
    std::vector<int32_t> last_index(n, -1);
    uint32_t n_big = 0, n_global = 0;
    const uint64_t scr_samples = (uint64_t)(ctts::SCR_WORDS - 4) / 2 * 32;   // region length the shared trim mask covers
    uint64_t big_region_max = 0;
    std::vector<uint32_t> row(n);
    for (uint32_t k = 0; k < max_tasks_per_utt; k++) {
        // inside a row: longest task first (it is the one a successor may have to wait for, and
        // longest-first balances the tail of the launch)
        uint32_t m = 0;
        for (uint32_t i = 0; i < n; i++)
            if (k < utt_tasks[order[i]].size()) row[m++] = order[i];
        std::stable_sort(row.begin(), row.begin() + m,
                         [&](uint32_t a, uint32_t b) { return utt_tasks[a][k].bound > utt_tasks[b][k].bound; });
        for (uint32_t i = 0; i < m; i++) {
            const uint32_t u = row[i];
            const HostTask& h = utt_tasks[u][k];
            ctts::RegionTask t{};
            t.utt = u;
            t.op_begin = h.op_begin;
            t.op_end = h.op_end;
            t.bound = (uint32_t)std::min<uint64_t>(h.bound, 0xffffffffull);
            t.pred = last_index[u];
            t.flags = (k + 1 == utt_tasks[u].size() ? ctts::TASK_LAST : 0u) | (to_pre[u] ? ctts::TASK_TO_PRE : 0u);
            if (h.bound > wcap) { t.flags |= ctts::TASK_GLOBAL; n_global++; }
            t.dst_cap = slot_cap[u];
            t.dst_off = slot_off[u];
            t.big = 0xffffffffu;
            if (h.region_max > scr_samples) {
                t.big = n_big++;
                big_region_max = std::max<uint64_t>(big_region_max, h.region_max);
            }
            last_index[u] = (int32_t)tasks.size();
            tasks.push_back(t);
        }
    }
    p->n_tasks = (uint32_t)tasks.size();
    p->n_global_tasks = n_global;
    if (n_big) p->trim_words = (uint32_t)(2 * ((big_region_max + 31) / 32) + 8);

#define CUP(call)                                                                               \
    do {                                                                                        \
        cudaError_t e_ = (call);                                                                \
        if (e_ != cudaSuccess) {                                                                \
            ctts_gpu_plan_destroy(p);                                                           \
            return fail(ctx, e_ == cudaErrorMemoryAllocation ? CTTS_GPU_ERR_OUT_OF_MEMORY : CTTS_GPU_ERR_CUDA, \
                        "%s: %s", #call, cudaGetErrorString(e_));                               \
        }                                                                                       \
    } while (0)
    auto up = [&](auto** d, const auto& h) -> cudaError_t {
        using T = typename std::remove_reference<decltype(h[0])>::type;
        cudaError_t e = cudaMalloc(reinterpret_cast<void**>(d), std::max<size_t>(h.size(), 1) * sizeof(T));
        if (e != cudaSuccess || h.empty()) return e;
        return cudaMemcpyAsync(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream);
    };
    CUP(up(&p->d_ops, ops));
    CUP(up(&p->d_tasks, tasks));
    CUP(up(&p->d_stasks, stasks));
    CUP(up(&p->d_ola_task, ola_task));
    CUP(up(&p->d_ola_first, ola_first));
    CUP(cudaMalloc(reinterpret_cast<void**>(&p->d_counts), std::max<size_t>(n, 1) * 4));
    CUP(cudaMalloc(reinterpret_cast<void**>(&p->d_pre_counts), std::max<size_t>(n, 1) * 4));
    CUP(cudaMalloc(reinterpret_cast<void**>(&p->d_err), std::max<size_t>(n, 1) * 4));
    if (p->trim_words) CUP(cudaMalloc(reinterpret_cast<void**>(&p->d_trim), (size_t)n_big * p->trim_words * 4));
    CUP(cudaMalloc(reinterpret_cast<void**>(&p->d_chain), std::max<size_t>(tasks.size(), 1) * 8));
    CUP(cudaMemsetAsync(p->d_chain, 0, std::max<size_t>(tasks.size(), 1) * 8, ctx->stream));
    CUP(cudaMalloc(reinterpret_cast<void**>(&p->d_ticket), 4));
    if (p->n_stretch) {
        CUP(cudaMalloc(reinterpret_cast<void**>(&p->d_pre), pre_total * sizeof(int16_t)));
        CUP(cudaMalloc(reinterpret_cast<void**>(&p->d_frame_pos), std::max<uint64_t>(pos_total, 1) * 4));
        CUP(cudaMalloc(reinterpret_cast<void**>(&p->d_n_frames), (size_t)p->n_stretch * 4));
    }
    CUP(cudaFuncSetAttribute(ctts::assemble_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem_bytes));
    // the host vectors above must outlive the async copies
    CUP(cudaStreamSynchronize(ctx->stream));
#undef CUP

    p->info.kernel_launches = 1 + (p->n_stretch ? 2 : 0);
    p->info.n_stretch = p->n_stretch;
    p->info.gather_samples = gather;
    p->info.bound_samples = std::accumulate(bound.begin(), bound.end(), (uint64_t)0);
    p->info.smem_bytes = p->smem_bytes;
    p->info.window_samples = wcap;
    p->info.halo_samples = hcap;
    p->info.threads = ctts::ASM_THREADS;
    {
        int occ = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ctts::assemble_kernel, ctts::ASM_THREADS, p->smem_bytes) != cudaSuccess || occ < 1)
            occ = 1;
        p->grid = (uint32_t)std::min<uint64_t>((uint64_t)occ * (uint64_t)ctx->sm_count, std::max<uint32_t>(p->n_tasks, 1));
        p->info.n_tasks = p->n_tasks;
        p->info.n_global_tasks = p->n_global_tasks;
        p->info.ctas_per_sm = (uint32_t)occ;
        p->info.grid = p->grid;
    }
    *out = p;
    return CTTS_GPU_OK;
}

uint64_t ctts_gpu_plan_out_samples(const ctts_gpu_plan* p) { return p ? p->offsets[p->n_utts] : 0; }

int ctts_gpu_plan_out_offsets(const ctts_gpu_plan* p, uint64_t* offsets) {
    if (!p || !offsets) return CTTS_GPU_ERR_INVALID_ARG;
    std::copy(p->offsets.begin(), p->offsets.end(), offsets);
    return CTTS_GPU_OK;
}

int ctts_gpu_plan_info(const ctts_gpu_plan* p, ctts_gpu_run_info* info) {
    if (!p || !info) return CTTS_GPU_ERR_INVALID_ARG;
    *info = p->info;
    return CTTS_GPU_OK;
}

int ctts_gpu_plan_run(ctts_gpu_ctx* ctx, ctts_gpu_plan* p, int16_t* d_pcm_out) {
    if (!ctx || !p || p->ctx != ctx) return CTTS_GPU_ERR_INVALID_ARG;
    CU(ctx, cudaSetDevice(ctx->device));
    if (!d_pcm_out) {
        if (!p->d_out_owned)
            CU(ctx, cudaMalloc(reinterpret_cast<void**>(&p->d_out_owned),
                               std::max<uint64_t>(p->offsets[p->n_utts], 8) * sizeof(int16_t)));
        d_pcm_out = p->d_out_owned;
    }
    p->d_out_last = d_pcm_out;
    if (p->n_utts == 0) return CTTS_GPU_OK;
    cudaStream_t st = ctx->stream;
    CU(ctx, cudaMemsetAsync(p->d_counts, 0, (size_t)p->n_utts * 4, st));
    CU(ctx, cudaMemsetAsync(p->d_pre_counts, 0, (size_t)p->n_utts * 4, st));
    CU(ctx, cudaMemsetAsync(p->d_err, 0, (size_t)p->n_utts * 4, st));
    CU(ctx, cudaMemsetAsync(p->d_ticket, 0, 4, st));
    p->epoch++;
    if (p->epoch == 0) p->epoch = 1;   // (2^32 runs later) the chain words of the last lap are long gone

    if (p->n_tasks) {
        ctts::AsmArgs a{};
        a.pool = ctx->d_pool;
        a.unit_off = ctx->d_unit_off;
        a.unit_cnt = ctx->d_unit_cnt;
        a.n_units = ctx->n_units;
        a.tab.fade_out = ctx->d_tables;
        a.tab.fade_in = ctx->d_tables + 1024;
        a.tab.sine = ctx->d_tables + 2048;
        a.tab.hann256 = ctx->d_tables + 3072;
        a.tab.hann512 = ctx->d_tables + 3328;
        a.tab.xfade4 = reinterpret_cast<const float4*>(ctx->d_tables + 3840);
        a.ops = p->d_ops;
        a.tasks = p->d_tasks;
        a.n_tasks = p->n_tasks;
        a.dst_final = d_pcm_out;
        a.dst_pre = p->d_pre;
        a.out_counts = p->d_counts;
        a.pre_counts = p->d_pre_counts;
        a.err = p->d_err;
        a.trim_scratch = p->d_trim;
        a.trim_scratch_words = p->trim_words;
        a.chain = p->d_chain;
        a.ticket = p->d_ticket;
        a.epoch = p->epoch;
        a.prm = p->prm;
        a.wcap = p->wcap;
        a.hcap = p->hcap;
        ctts::assemble_kernel<<<p->grid, ctts::ASM_THREADS, p->smem_bytes, st>>>(a);
        CU(ctx, cudaGetLastError());
    }

    if (p->n_stretch) {
        ctts::WsolaArgs w{};
        w.tasks = p->d_stasks;
        w.n_tasks = p->n_stretch;
        w.pre = p->d_pre;
        w.pre_counts = p->d_pre_counts;
        w.out = d_pcm_out;
        w.out_counts = p->d_counts;
        w.frame_pos = p->d_frame_pos;
        w.n_frames = p->d_n_frames;
        w.hann512 = ctx->d_tables + 3328;
        w.ola_block_task = p->d_ola_task;
        w.ola_block_first = p->d_ola_first;
        ctts::wsola_search_kernel<<<p->n_stretch, ctts::WS_THREADS, 0, st>>>(w);
        CU(ctx, cudaGetLastError());
        ctts::wsola_ola_kernel<<<p->n_ola_blocks, ctts::OLA_THREADS, 0, st>>>(w);
        CU(ctx, cudaGetLastError());
    }
    return CTTS_GPU_OK;
}

int ctts_gpu_plan_read_counts(ctts_gpu_ctx* ctx, ctts_gpu_plan* p, uint32_t* out_counts) {
    if (!ctx || !p || !out_counts) return CTTS_GPU_ERR_INVALID_ARG;
    CU(ctx, cudaSetDevice(ctx->device));
    if (p->n_utts == 0) return CTTS_GPU_OK;
    std::vector<uint32_t> err(p->n_utts);
    CU(ctx, cudaMemcpyAsync(out_counts, p->d_counts, (size_t)p->n_utts * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaMemcpyAsync(err.data(), p->d_err, (size_t)p->n_utts * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    for (uint32_t u = 0; u < p->n_utts; u++)
        if (err[u]) return fail(ctx, CTTS_GPU_ERR_DEVICE, "utterance %u: device error %u", u, err[u]);
    return CTTS_GPU_OK;
}

int ctts_gpu_plan_read_pcm(ctts_gpu_ctx* ctx, ctts_gpu_plan* p, int16_t* dst, uint64_t first, uint64_t n) {
    if (!ctx || !p || !dst || !p->d_out_last || first + n > p->offsets[p->n_utts]) return CTTS_GPU_ERR_INVALID_ARG;
    CU(ctx, cudaSetDevice(ctx->device));
    CU(ctx, cudaMemcpyAsync(dst, p->d_out_last + first, n * sizeof(int16_t), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    return CTTS_GPU_OK;
}

int ctts_gpu_plan_read_pre(ctts_gpu_ctx* ctx, ctts_gpu_plan* p, uint32_t u, int16_t* dst, uint64_t cap, uint64_t* n) {
    if (!ctx || !p || !dst || !n || u >= p->n_utts || p->pre_off[u] == ~0ull) return CTTS_GPU_ERR_INVALID_ARG;
    CU(ctx, cudaSetDevice(ctx->device));
    uint32_t cnt = 0;
    CU(ctx, cudaMemcpyAsync(&cnt, p->d_pre_counts + u, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    *n = cnt;
    uint64_t take = std::min<uint64_t>(cnt, cap);
    CU(ctx, cudaMemcpy(dst, p->d_pre + p->pre_off[u], take * sizeof(int16_t), cudaMemcpyDeviceToHost));
    return CTTS_GPU_OK;
}

int ctts_gpu_synth_batch(ctts_gpu_ctx* ctx, const ctts_batch_plan* plan, const ctts_assembly_params* params,
                         int16_t* pcm_out, const uint64_t* out_offsets, uint32_t* out_counts) {
    if (!ctx || !plan || !params || !pcm_out || !out_offsets || !out_counts) return CTTS_GPU_ERR_INVALID_ARG;
    ctts_gpu_plan* p = nullptr;
    int rc = ctts_gpu_plan_create(ctx, plan, params, out_offsets, &p);
    if (rc) return rc;
    uint64_t total = p->offsets[p->n_utts];
    if (total > ctx->batch_out_cap) {
        cudaFree(ctx->d_batch_out);
        ctx->d_batch_out = nullptr;
        ctx->batch_out_cap = 0;
        cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&ctx->d_batch_out), std::max<uint64_t>(total, 8) * sizeof(int16_t));
        if (e != cudaSuccess) {
            ctts_gpu_plan_destroy(p);
            return fail(ctx, CTTS_GPU_ERR_OUT_OF_MEMORY, "output buffer: %s", cudaGetErrorString(e));
        }
        ctx->batch_out_cap = total;
    }
    rc = ctts_gpu_plan_run(ctx, p, ctx->d_batch_out);
    if (!rc) rc = ctts_gpu_plan_read_counts(ctx, p, out_counts);
    if (!rc && total) {
        // one copy of the occupied span; slots keep their caller-chosen offsets
        uint64_t lo = out_offsets[0];
        cudaError_t e = cudaMemcpyAsync(pcm_out + lo, ctx->d_batch_out + lo, (total - lo) * sizeof(int16_t),
                                        cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) rc = fail(ctx, CTTS_GPU_ERR_CUDA, "D2H: %s", cudaGetErrorString(e));
    }
    ctts_gpu_plan_destroy(p);
    return rc;
}

}  // extern "C"
