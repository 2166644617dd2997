#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <time.h>
#include "p264b200_host.h"
int main(int argc, char **argv) {
    FILE *f = fopen(argv[1], "rb"); fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    uint8_t *d = malloc(n + 16), *pl = malloc(n + 16); fread(d, 1, n, f); fclose(f);
    for (int rep = 0; rep < 3; rep++) {
        p264b200_parser *p = p264b200_parser_open(argc > 2 ? atoi(argv[2]) : 0, 0);
        struct timespec t0, t1; clock_gettime(CLOCK_MONOTONIC, &t0);
        size_t pos = 0, start, len; int frames = 0;
        while (p264b200_annexb_next(d, n, &pos, &start, &len)) {
            int type, ref; int l = p264b200_nal_unescape(d + start, (int)len, pl, &type, &ref);
            p264b200_frame_syntax fs; int got = 0;
            p264b200_parser_nal(p, type, ref, pl, l, &fs, &got);
            frames += got;
        }
        clock_gettime(CLOCK_MONOTONIC, &t1);
        double s = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
        printf("%d frames in %.3f s: %.3f ms per picture\n", frames, s, 1e3 * s / frames);
        p264b200_parser_close(p);
    }
    return 0;
}
