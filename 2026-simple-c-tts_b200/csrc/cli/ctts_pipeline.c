/*
 * ctts_pipeline.c -- ctts_b200_synth_texts (include/ctts_b200.h): the text front end pipelined into the GPU
 * back end.  Plain C + pthreads over the two C-ABI libraries.
 *
 * The reference's ctts_synthesize (ctts.c:3623) does, per utterance, normalisation (:3638-3655), the walk
 * with unit selection (:3689-3871, :1406) and every sample loop in one call on one core.  Here:
 *   planner threads   take small groups of utterances (16) off a shared counter and plan them with the unchanged
 *                     front end (ctts_front_plan_batch_threads(..., 1, ...): re-entrant, the handle is only
 *                     read), at most LOOKAHEAD groups ahead of the device;
 *   the calling thread takes, in order, every plan that is ready -- up to piece_utts utterances -- joins them
 *                     into one piece and submits it to a ctts_gpu_session (asynchronous: several pieces are
 *                     compiled / assembled / copied at any time).  So the first piece reaches the device after
 *                     ~1.5 ms of planning (a group is planned by one thread at ~100 us per utterance, while
 *                     the device->host copy alone consumes an utterance every ~25 us), and once the planners
 *                     are ahead the pieces are large enough to keep the kernels and the copies efficient.
 * Nothing here touches a sample.
 */
#define _POSIX_C_SOURCE 200809L
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include "ctts_b200.h"

#define LOOKAHEAD 96   /* groups planned ahead of the one being submitted (bounds the memory held in plans) */
#define GROUP_UTTS 16  /* utterances per planner job */

typedef struct {
    ctts_batch_plan plan;
    int state;   /* 0: not planned, 1: planned, <0: error code */
} piece_slot;

typedef struct {
    ctts_front* front;
    const char* const* texts;
    const float* speeds;
    uint32_t* stats;
    uint32_t n, n_pieces;
    uint32_t* piece_begin;     /* n_pieces + 1: the first pieces are small so that the device starts early */
    piece_slot* slots;
    pthread_mutex_t mu;
    pthread_cond_t cv_ready;   /* a piece became ready */
    pthread_cond_t cv_room;    /* the submitter moved on */
    uint32_t next;             /* next piece to plan */
    uint32_t consumed;         /* pieces the submitter is done with */
    int stop;
    double t0, first_plan_s, all_plans_s;
    uint32_t planned;
} pipeline;

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static void* planner(void* arg) {
    pipeline* P = arg;
    for (;;) {
        pthread_mutex_lock(&P->mu);
        while (!P->stop && P->next < P->n_pieces && P->next >= P->consumed + LOOKAHEAD) pthread_cond_wait(&P->cv_room, &P->mu);
        if (P->stop || P->next >= P->n_pieces) {
            pthread_mutex_unlock(&P->mu);
            return NULL;
        }
        const uint32_t i = P->next++;
        pthread_mutex_unlock(&P->mu);
        const uint32_t u0 = P->piece_begin[i], cnt = P->piece_begin[i + 1] - u0;
        ctts_batch_plan pl;
        int rc = ctts_front_plan_batch_threads(P->front, P->texts + u0, P->speeds ? P->speeds + u0 : NULL, cnt, 1, &pl,
                                               P->stats ? P->stats + 2 * (size_t)u0 : NULL);
        pthread_mutex_lock(&P->mu);
        if (rc == 0) {
            P->slots[i].plan = pl;
            P->slots[i].state = 1;
        } else {
            P->slots[i].state = rc < 0 ? rc : -1;
        }
        const double t = now_s() - P->t0;
        if (P->planned++ == 0) P->first_plan_s = t;
        P->all_plans_s = t;
        pthread_cond_broadcast(&P->cv_ready);
        pthread_mutex_unlock(&P->mu);
    }
}

uint64_t ctts_b200_capacity_hint(ctts_front* front, const char* const* texts, const float* speeds, uint32_t n) {
    if (!front || (n && !texts)) return 0;
    /* every character yields at most one unit or one pause; pauses are shorter than the longest unit for
     * any sane config, but take the larger of the two anyway */
    uint64_t per_char = ctts_front_max_unit_samples(front);
    if (per_char < 8192) per_char = 8192;
    uint64_t total = 0;
    for (uint32_t u = 0; u < n; u++) {
        uint64_t chars = 8;
        for (const unsigned char* c = (const unsigned char*)(texts[u] ? texts[u] : ""); *c; c++)
            chars += (*c >= '0' && *c <= '9') ? 16 : 1;   /* a digit expands to words ("novecentos e ") */
        uint64_t s = chars * per_char;
        if (speeds && speeds[u] < 1.0f) s *= 2;   /* hop = 128 / speed <= 256 */
        total += s + 1024;
    }
    return total;
}

int ctts_b200_synth_texts(ctts_front* front, ctts_gpu_ctx* gpu, const char* const* texts, const float* speeds,
                          uint32_t n, int16_t* pcm_out, uint64_t capacity, uint64_t* out_offsets,
                          uint32_t* out_counts, uint32_t* stats, uint64_t* samples_used,
                          const ctts_b200_options* opt, ctts_b200_timing* timing) {
    if (!front || !gpu || (n && (!texts || !out_offsets || !out_counts))) return CTTS_GPU_ERR_INVALID_ARG;
    pipeline P;
    memset(&P, 0, sizeof P);
    P.front = front;
    P.texts = texts;
    P.speeds = speeds;
    P.stats = stats;
    P.n = n;
    const uint32_t piece_utts = opt && opt->piece_utts ? opt->piece_utts : 128;
    const uint32_t group = piece_utts < GROUP_UTTS ? piece_utts : GROUP_UTTS;
    long cores = sysconf(_SC_NPROCESSORS_ONLN);
    uint32_t T = opt && opt->threads ? opt->threads : (uint32_t)(cores > 1 ? cores - 1 : 1);
    if (T > 32) T = 32;
    if (T < 1) T = 1;
    P.n_pieces = (n + group - 1) / group;
    P.piece_begin = malloc(((size_t)P.n_pieces + 1) * sizeof *P.piece_begin);
    if (!P.piece_begin) return CTTS_GPU_ERR_OUT_OF_MEMORY;
    for (uint32_t i = 0; i <= P.n_pieces; i++) P.piece_begin[i] = (uint64_t)i * group < n ? i * group : n;
    P.t0 = now_s();
    if (T > P.n_pieces) T = P.n_pieces;
    if (T < 1) T = 1;

    ctts_assembly_params prm;
    ctts_front_params(front, &prm);
    ctts_gpu_session* ses = NULL;
    int rc = ctts_gpu_session_begin(gpu, &prm, pcm_out, capacity, opt ? opt->on_piece : NULL, opt ? opt->user : NULL, &ses);
    if (rc) {
        free(P.piece_begin);
        return rc;
    }
    P.slots = calloc(P.n_pieces ? P.n_pieces : 1, sizeof *P.slots);
    pthread_t* tids = calloc(T, sizeof *tids);
    if (!P.slots || !tids) {
        free(P.slots);
        free(tids);
        free(P.piece_begin);
        ctts_gpu_session_end(ses, NULL);
        return CTTS_GPU_ERR_OUT_OF_MEMORY;
    }
    pthread_mutex_init(&P.mu, NULL);
    pthread_cond_init(&P.cv_ready, NULL);
    pthread_cond_init(&P.cv_room, NULL);
    uint32_t started = 0;
    for (; started < T; started++)
        if (pthread_create(&tids[started], NULL, planner, &P) != 0) break;
    double waited = 0.0, all_submitted = 0.0;
    uint32_t n_submitted = 0;
    if (started == 0 && P.n_pieces) rc = CTTS_GPU_ERR_OUT_OF_MEMORY;
    ctts_batch_plan joined;                 /* the piece being submitted when it spans several plans */
    uint32_t* j_begin = NULL;
    float* j_speed = NULL;
    ctts_plan_op* j_ops = NULL;
    size_t j_utts_cap = 0, j_ops_cap = 0;
    for (uint32_t i = 0; i < P.n_pieces && !rc;) {
        const double w0 = now_s();
        pthread_mutex_lock(&P.mu);
        while (P.slots[i].state == 0) pthread_cond_wait(&P.cv_ready, &P.mu);
        /* every further plan that is ready right now, up to piece_utts utterances */
        uint32_t j = i + 1;
        while (j < P.n_pieces && P.slots[j].state != 0 && P.piece_begin[j + 1] - P.piece_begin[i] <= piece_utts) j++;
        int st = 1;
        for (uint32_t k = i; k < j; k++)
            if (P.slots[k].state < 0) st = P.slots[k].state;
        pthread_mutex_unlock(&P.mu);
        waited += now_s() - w0;
        if (st < 0) {
            rc = st;
            break;
        }
        const uint32_t u0 = P.piece_begin[i];
        const ctts_batch_plan* piece = &P.slots[i].plan;
        if (j > i + 1) {
            size_t nu = 0, no = 0;
            for (uint32_t k = i; k < j; k++) {
                nu += P.slots[k].plan.n_utts;
                no += P.slots[k].plan.n_ops;
            }
            if (nu + 1 > j_utts_cap) {
                j_utts_cap = 2 * (nu + 1);
                free(j_begin);
                free(j_speed);
                j_begin = malloc(j_utts_cap * sizeof *j_begin);
                j_speed = malloc(j_utts_cap * sizeof *j_speed);
            }
            if (no + 1 > j_ops_cap) {
                j_ops_cap = 2 * (no + 1);
                free(j_ops);
                j_ops = malloc(j_ops_cap * sizeof *j_ops);
            }
            if (!j_begin || !j_speed || !j_ops || no > 0xffffffffu) {
                rc = CTTS_GPU_ERR_OUT_OF_MEMORY;
                break;
            }
            uint32_t au = 0, ao = 0;
            for (uint32_t k = i; k < j; k++) {
                const ctts_batch_plan* q = &P.slots[k].plan;
                for (uint32_t u = 0; u < q->n_utts; u++) {
                    j_begin[au + u] = ao + q->utt_op_begin[u];
                    j_speed[au + u] = q->speed[u];
                }
                if (q->n_ops) memcpy(j_ops + ao, q->ops, (size_t)q->n_ops * sizeof *j_ops);
                au += q->n_utts;
                ao += q->n_ops;
            }
            j_begin[au] = ao;
            joined.n_utts = au;
            joined.n_ops = ao;
            joined.utt_op_begin = j_begin;
            joined.speed = j_speed;
            joined.ops = j_ops;
            piece = &joined;
        }
        rc = ctts_gpu_session_submit(ses, piece, out_offsets + u0, out_counts + u0);
        for (uint32_t k = i; k < j; k++) {
            ctts_front_plan_free(&P.slots[k].plan);   /* the session keeps nothing of the plan */
            P.slots[k].state = 2;
        }
        pthread_mutex_lock(&P.mu);
        P.consumed = j;
        pthread_cond_broadcast(&P.cv_room);
        pthread_mutex_unlock(&P.mu);
        all_submitted = now_s() - P.t0;
        n_submitted++;
        i = j;
    }
    free(j_begin);
    free(j_speed);
    free(j_ops);
    pthread_mutex_lock(&P.mu);
    P.stop = 1;
    pthread_cond_broadcast(&P.cv_room);
    pthread_mutex_unlock(&P.mu);
    for (uint32_t t = 0; t < started; t++) pthread_join(tids[t], NULL);
    for (uint32_t i = 0; i < P.n_pieces; i++)
        if (P.slots[i].state == 1) ctts_front_plan_free(&P.slots[i].plan);
    const int rc_end = ctts_gpu_session_end(ses, samples_used);
    if (!rc) rc = rc_end;
    if (timing) {
        timing->first_plan_s = P.first_plan_s;
        timing->all_plans_s = P.all_plans_s;
        timing->all_submitted_s = all_submitted;
        timing->done_s = now_s() - P.t0;
        timing->wait_for_plans_s = waited;
        timing->pieces = n_submitted;
    }
    pthread_mutex_destroy(&P.mu);
    pthread_cond_destroy(&P.cv_ready);
    pthread_cond_destroy(&P.cv_room);
    free(P.slots);
    free(tids);
    free(P.piece_begin);
    return rc;
}
