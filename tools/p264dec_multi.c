/*
 * p264dec_multi -- decode N H.264 Annex-B streams concurrently on one GPU:
 *     p264dec_multi [-t threads] [-o out_prefix] <in1.264> [in2.264 ...]
 *     p264dec_multi [-t threads] [-o out_prefix] -n N <in.264>        (N copies of one stream)
 * One lane per stream (include/p264b200_host.h, p264b200_multi_*); with -o every stream s is written
 * as <out_prefix><s>.yuv (tight I420, the format the reference CLI writes, p264decoder.c:126-156).
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "p264b200_host.h"

static uint8_t *read_file(const char *path, size_t *size)
{
    FILE *f = fopen(path, "rb");
    if (!f) {
        fprintf(stderr, "open h264 stream file: %s failed\n", path);
        return NULL;
    }
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    uint8_t *d = malloc(n + 16);
    if (!d || fread(d, 1, n, f) != (size_t)n) {
        fclose(f);
        free(d);
        return NULL;
    }
    fclose(f);
    *size = (size_t)n;
    return d;
}

int main(int argc, char **argv)
{
    int threads = 0, copies = 0, i = 1;
    const char *prefix = NULL;
    while (i < argc && argv[i][0] == '-' && argv[i][1]) {
        if (!strcmp(argv[i], "-t") && i + 1 < argc)
            threads = atoi(argv[i + 1]), i += 2;
        else if (!strcmp(argv[i], "-n") && i + 1 < argc)
            copies = atoi(argv[i + 1]), i += 2;
        else if (!strcmp(argv[i], "-o") && i + 1 < argc)
            prefix = argv[i + 1], i += 2;
        else
            break;
    }
    if (i >= argc) {
        fprintf(stderr, "p264 multi-stream decoder (B200):\n\n      [-t threads] [-o out_prefix] <in1.264> [in2.264 ...]\n"
                        "      [-t threads] [-o out_prefix] -n N <in.264>\n");
        return -1;
    }
    const int n = copies > 0 ? copies : argc - i;
    if (n < 1 || n > 256) {
        fprintf(stderr, "1..256 streams\n");
        return -1;
    }
    uint8_t **data = calloc(n, sizeof(*data));
    size_t *size = calloc(n, sizeof(*size));
    FILE **out = calloc(n, sizeof(*out));
    for (int s = 0; s < n; s++) {
        if (copies > 0 && s > 0) {
            data[s] = data[0], size[s] = size[0];
        } else if (!(data[s] = read_file(argv[i + (copies > 0 ? 0 : s)], &size[s])))
            return -1;
        if (prefix) {
            char name[1024];
            snprintf(name, sizeof(name), "%s%d.yuv", prefix, s);
            if (!(out[s] = fopen(name, "wb"))) {
                fprintf(stderr, "cannot create %s\n", name);
                return -1;
            }
        }
    }
    p264b200_multi_cfg cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.n_streams = n;
    cfg.n_threads = threads;
    p264b200_multi *m = NULL;
    if (p264b200_multi_open(&m, &cfg) < 0) return -1;
    for (int s = 0; s < n; s++) p264b200_multi_set_stream(m, s, data[s], size[s]);
    fprintf(stderr, "decoding start... (%d streams)\n", n);
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    long frames = 0;
    int r;
    long first = 0;
    struct timespec tf = t0;
    while ((r = p264b200_multi_step(m, NULL)) > 0) {
        if (!frames) {
            first = r;   /* the first step creates the CUDA context and the engine */
            clock_gettime(CLOCK_MONOTONIC, &tf);
        }
        frames += r;
        if (prefix)
            for (int s = 0; s < n; s++) {
                int w, h;
                const uint8_t *pic = p264b200_multi_picture(m, s, &w, &h);
                if (pic) fwrite(pic, 1, (size_t)w * h * 3 / 2, out[s]);
            }
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (r < 0) fprintf(stderr, "decode error %d\n", r);
    const double secs = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
    fprintf(stderr, "decoded total %ld frames \n", frames);
    fprintf(stderr, "decoding speed: %.2f fps\n", frames / secs);
    const double steady = (t1.tv_sec - tf.tv_sec) + 1e-9 * (t1.tv_nsec - tf.tv_nsec);
    if (frames > first && steady > 0) fprintf(stderr, "after the first step (CUDA context + engine creation): %.2f fps\n", (frames - first) / steady);
    p264b200_multi_close(m);
    for (int s = 0; s < n; s++)
        if (out[s]) fclose(out[s]);
    return r < 0 ? -1 : 0;
}
