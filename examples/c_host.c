/* A host written in plain C against include/afp.h - the same calls a cgo / JNI / Rust-FFI binding would
 * make.  It runs the body of the reference's audio_callback (stream_process.py:79-120) for a few blocks of
 * a stereo session on the GPU: host buffers in, host buffers out, state carried between blocks, and checks
 * the result against a direct-form convolution computed here in double precision.
 *
 *   gcc -std=c99 -Wall -Iinclude -o /tmp/c_host examples/c_host.c \
 *       -Laudio_filtering_pipeline_b200 -lafp_b200 -lm -Wl,-rpath,$PWD/audio_filtering_pipeline_b200
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "afp.h"

#define CHECK(call) do { int rc_ = (call); if (rc_ != AFP_OK) { \
    fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, afp_last_error(ctx)); return 1; } } while (0)

int main(void) {
    enum { S = 2, FRAMES = 512, BLOCKS = 4, NTAPS = 101 };
    afp_ctx* ctx = NULL;
    afp_config cfg = {0};
    cfg.device = 0; cfg.n_streams = S; cfg.max_frames = FRAMES; cfg.upsample = 1;
    cfg.chain = AFP_CHAIN_DAC; cfg.out_format = AFP_FMT_F32; cfg.dither = AFP_DITHER_NONE; cfg.eq_mode = AFP_EQ_NONE;
    printf("%s\n", afp_version());
    CHECK(afp_create(&cfg, &ctx));

    double taps[NTAPS];                       /* windowed-sinc low-pass, the job of create_fir_filter */
    const double pi = 3.14159265358979323846;
    double sum = 0.0;
    for (int k = 0; k < NTAPS; ++k) {
        const double m = k - (NTAPS - 1) / 2.0, fc = 0.2;
        const double sinc = m == 0.0 ? 2.0 * fc : sin(2.0 * pi * fc * m) / (pi * m);
        taps[k] = sinc * (0.54 - 0.46 * cos(2.0 * pi * k / (NTAPS - 1)));
        sum += taps[k];
    }
    for (int k = 0; k < NTAPS; ++k) taps[k] /= sum;
    CHECK(afp_set_taps(ctx, taps, NTAPS));

    float *in = NULL, *out = NULL;            /* pinned staging: in[S][FRAMES], out[S][FRAMES] per block */
    CHECK(afp_host_alloc((void**)&in, sizeof(float) * S * FRAMES));
    CHECK(afp_host_alloc((void**)&out, sizeof(float) * S * FRAMES));
    float* x = (float*)malloc(sizeof(float) * S * FRAMES * BLOCKS);
    float* y = (float*)malloc(sizeof(float) * S * FRAMES * BLOCKS);
    unsigned lcg = 12345u;
    for (int i = 0; i < S * FRAMES * BLOCKS; ++i) {
        lcg = lcg * 1664525u + 1013904223u;
        x[i] = 0.2f * ((float)(lcg >> 8) / 16777216.0f - 0.5f);      /* stream s, sample n at x[s*FRAMES*BLOCKS + n] */
    }
    for (int b = 0; b < BLOCKS; ++b) {        /* one "callback" per block */
        for (int s = 0; s < S; ++s)
            for (int i = 0; i < FRAMES; ++i) in[s * FRAMES + i] = x[s * FRAMES * BLOCKS + b * FRAMES + i];
        CHECK(afp_process_host(ctx, in, out, FRAMES));
        for (int s = 0; s < S; ++s)
            for (int i = 0; i < FRAMES; ++i) y[s * FRAMES * BLOCKS + b * FRAMES + i] = out[s * FRAMES + i];
    }
    double max_err = 0.0;
    for (int s = 0; s < S; ++s)
        for (int n = 0; n < FRAMES * BLOCKS; ++n) {
            double acc = 0.0;
            for (int k = 0; k < NTAPS && k <= n; ++k) acc += (double)(float)taps[k] * x[s * FRAMES * BLOCKS + n - k];
            const double e = fabs(acc - y[s * FRAMES * BLOCKS + n]);
            if (e > max_err) max_err = e;
        }
    afp_stage_info info;
    CHECK(afp_get_stage_info(ctx, 0, &info));
    printf("c_host: %d streams x %d blocks of %d frames, kernel %d, max |err| = %.3g -> %s\n", S, BLOCKS, FRAMES,
           info.kernel, max_err, max_err < 1e-5 ? "ok" : "FAIL");
    free(x); free(y);
    afp_host_free(in); afp_host_free(out);
    afp_destroy(ctx);
    return max_err < 1e-5 ? 0 : 2;
}
