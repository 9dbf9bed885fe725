/* Minimal C client of the rnnoise_* ABI (same call pattern as the reference's C demo client:
 * create(NULL), process 480-sample frames IN PLACE, round to int16, drop the first frame).
 * usage: demo_client <in.raw> <out.raw>   (16-bit LE mono 48 kHz) */
#include <math.h>
#include <stdio.h>
#include "rnnoise.h"

int main(int argc, char **argv) {
    if (argc != 3) {
        fprintf(stderr, "usage: %s <noisy.raw> <denoised.raw>\n", argv[0]);
        return 2;
    }
    const int n = rnnoise_get_frame_size();
    if (n != 480) return 3;
    DenoiseState *st = rnnoise_create(NULL);
    if (!st) {
        fprintf(stderr, "rnnoise_create failed: %s\n", rnnoise_last_error());
        return 4;
    }
    FILE *fi = fopen(argv[1], "rb"), *fo = fopen(argv[2], "wb");
    if (!fi || !fo) return 5;
    float x[480];
    short pcm[480];
    int first = 1;
    double vad_sum = 0.0;
    while (fread(pcm, sizeof(short), 480, fi) == 480) {
        for (int i = 0; i < 480; i++) x[i] = pcm[i];
        vad_sum += rnnoise_process_frame(st, x, x);
        for (int i = 0; i < 480; i++) pcm[i] = (short)roundf(x[i]);
        if (!first) fwrite(pcm, sizeof(short), 480, fo);
        first = 0;
    }
    rnnoise_destroy(st);
    fclose(fi);
    fclose(fo);
    printf("vad_sum %.6f\n", vad_sum);
    return 0;
}
