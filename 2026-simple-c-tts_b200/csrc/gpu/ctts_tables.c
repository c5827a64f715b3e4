/*
 * ctts_tables.c -- host-computed lookup tables uploaded by ctts_gpu_init.
 *
 * Compiled as C99 with -ffp-contract=off (not by nvcc) on purpose: the
 * reference fills these tables with glibc cosf/sinf under C promotion rules
 * (PI is a double literal, ctts.c:45), and CUDA's cosf is not bit-identical to
 * libm's.  Bit-exact PCM needs bit-exact tables, so they are produced here the
 * way ctts.c:60-73 (fade LUTs), :2198-2204 (256-sample Hann) and :1624
 * (512-sample WSOLA Hann) produce them.
 */
#include <math.h>
#include <stddef.h>

#define CTTS_PI 3.14159265358979323846

void ctts_host_tables(float* fade_out, float* fade_in, float* sine, float* hann256, float* hann512) {
    for (int i = 0; i < 1024; i++) {
        float t = (float)i / (float)(1024 - 1);
        fade_out[i] = 0.5f * (1.0f + cosf(CTTS_PI * t));
        fade_in[i] = 0.5f * (1.0f - cosf(CTTS_PI * t));
        sine[i] = sinf(t * CTTS_PI * 0.5f);
    }
    for (int i = 0; i < 256; i++) hann256[i] = 0.5f * (1.0f - cosf(2.0f * CTTS_PI * i / 256));
    for (size_t i = 0; i < 512; i++)
        hann512[i] = 0.5f * (1.0f - cosf(2.0f * (float)CTTS_PI * (float)i / (float)512));
}
