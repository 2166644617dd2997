/*
 * p264dec_b200 -- command-line decoder with the reference CLI's grammar
 *     p264dec_b200 -d <in.264> [recon.yuv] [origin.yuv]
 * (p264decoder.c:69-94,164-381) written against the drop-in API of include/p264_b200.h.
 * Differences: the whole file is mapped instead of a 3 MB sliding buffer (no NAL size limit,
 * p264decoder.c:48), and the third argument is accepted and ignored like the reference does.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "p264_b200.h"
#include "p264b200_host.h"

static void write_plane(FILE *fp, const uint8_t *p, int stride, int w, int h)
{
    for (int y = 0; y < h; y++) fwrite(p + (size_t)y * stride, 1, w, fp);
}

int main(int argc, char **argv)
{
    if (argc < 3 || strcmp(argv[1], "-d")) {
        fprintf(stderr, "p264 Decoder (B200):\n\n      -d <test.264> [recon.yuv] [origin.yuv]\n");
        return -1;
    }
    FILE *fin = fopen(argv[2], "rb");
    if (!fin) {
        fprintf(stderr, "open h264 stream file: %s failed\n", argv[2]);
        return -1;
    }
    fseek(fin, 0, SEEK_END);
    long size = ftell(fin);
    fseek(fin, 0, SEEK_SET);
    uint8_t *data = malloc(size + 16);
    if (fread(data, 1, size, fin) != (size_t)size) return -1;
    fclose(fin);
    FILE *fout = argc >= 4 ? fopen(argv[3], "wb") : NULL;

    p264_param_t param;
    p264_param_default(&param);
    fprintf(stderr, "decoding start...\n");
    p264_t *h = p264_decoder_open(&param);
    if (!h) {
        fprintf(stderr, "p264_decoder_open failed\n");
        return -1;
    }
    p264_nal_t nal;
    nal.p_payload = malloc(size + 16);
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    size_t pos = 0, start, n;
    int frames = 0;
    while (p264b200_annexb_next(data, size, &pos, &start, &n)) {
        p264_picture_t *pic = NULL;
        p264_nal_decode(&nal, data + start, (int)n);
        p264_decoder_decode(h, &pic, &nal);
        if (pic) {
            frames++;
            if (fout) {
                write_plane(fout, pic->img.plane[0], pic->img.i_stride[0], pic->i_width, pic->i_height);
                write_plane(fout, pic->img.plane[1], pic->img.i_stride[1], pic->i_width >> 1, pic->i_height >> 1);
                write_plane(fout, pic->img.plane[2], pic->img.i_stride[2], pic->i_width >> 1, pic->i_height >> 1);
            }
        }
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (frames > 0) {
        double secs = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
        fprintf(stderr, "decoded total %d frames \n", frames);
        fprintf(stderr, "decoding speed: %.2f fps\n", frames / secs);
    }
    p264_decoder_close(h);
    if (fout) fclose(fout);
    free(nal.p_payload);
    free(data);
    return 0;
}
