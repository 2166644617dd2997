// Closed-GOP splitter and ordered merge: one long Annex-B stream is cut at its IDR pictures,
// the closed GOPs are dealt round-robin to the lanes of one batched engine (GOP g -> lane g mod L) and the pictures
// come back in stream order.  This is the only way ONE stream can use more than one lane: inside a GOP every picture
// depends on the previous one.  The reference decodes a stream strictly serially (decoder/decoder.c:745-806); the cut
// relies on what its NAL switch does at an IDR (decoder/decoder.c:43-64: the DPB is emptied), so a GOP decodes the
// same whether or not the pictures before it were decoded.
//
// Built on the multi-stream decoder (multi.cc): lane l is fed the byte stream  [parameter sets] GOP l, GOP l+L, ...
#include <cstdio>
#include <cstring>
#include <deque>
#include <new>
#include <vector>

#include "../../../include/p264b200_host.h"

namespace {
struct Gop {
    size_t begin, end;   // byte range [begin, end) of the stream, start codes included
    int pictures;        // slice NAL units (the reference decodes one slice per picture)
};

// NAL walk: byte ranges of the closed GOPs (each starts at the SPS / PPS / other non-slice units that directly
// precede its IDR slice) and of the parameter sets seen before the first IDR
void scan(const uint8_t *buf, size_t bytes, std::vector<Gop> &gops, std::vector<uint8_t> &params)
{
    size_t pos = 0, start, n;
    size_t pending = (size_t)-1;   // first non-slice NAL since the last slice: where the next GOP would begin
    size_t prev_end = 0;
    while (p264b200_annexb_next(buf, bytes, &pos, &start, &n)) {
        const int type = buf[start] & 0x1f;
        // the unit's own start code: the zero bytes + 0x01 in front of it (at most 4 bytes are claimed)
        size_t sc = start;
        if (sc >= 3 && buf[sc - 1] == 1 && buf[sc - 2] == 0 && buf[sc - 3] == 0) sc -= 3;
        if (sc > prev_end && buf[sc - 1] == 0) sc--;
        if (type == 5) {
            const size_t b = pending != (size_t)-1 ? pending : sc;
            if (!gops.empty()) gops.back().end = b;
            gops.push_back({b, bytes, 1});
            pending = (size_t)-1;
        } else if (type == 1) {
            if (!gops.empty()) gops.back().pictures++;
            pending = (size_t)-1;
        } else {
            if (pending == (size_t)-1) pending = sc;
            if ((type == 7 || type == 8)) {
                // every parameter set of the stream goes in front of every lane (a later GOP may rely on sets sent once
                // at the start); sets that are repeated in front of their IDR are simply parsed twice
                static const uint8_t code[4] = {0, 0, 0, 1};
                params.insert(params.end(), code, code + 4);
                params.insert(params.end(), buf + start, buf + start + n);
            }
        }
        prev_end = start + n;
    }
}
}  // namespace

struct p264b200_gopdec {
    p264b200_multi *multi = nullptr;
    int lanes = 0;
    std::vector<Gop> gops;
    std::vector<std::vector<uint8_t>> lane_stream;      // the byte stream fed to each lane
    std::vector<std::deque<std::vector<uint8_t>>> fifo;  // decoded pictures per lane, oldest first
    std::vector<uint8_t> produced, current;
    size_t gop_i = 0;      // next picture to deliver: GOP gop_i, picture pic_i
    int pic_i = 0;
    int width = 0, height = 0;
    bool drained = false;
};

extern "C" {

int p264b200_gop_scan(const uint8_t *annexb, size_t bytes, size_t *gop_begin, int32_t *gop_pictures, int max_gops)
{
    if (!annexb) return P264B200_EINVAL;
    std::vector<Gop> gops;
    std::vector<uint8_t> params;
    scan(annexb, bytes, gops, params);
    for (size_t i = 0; i < gops.size() && (int)i < max_gops; i++) {
        if (gop_begin) gop_begin[i] = gops[i].begin;
        if (gop_pictures) gop_pictures[i] = gops[i].pictures;
    }
    return (int)gops.size();
}

int p264b200_gopdec_open(p264b200_gopdec **out, int device, int lanes, int threads, const uint8_t *annexb, size_t bytes)
{
    if (!out || !annexb || lanes < 1 || lanes > 256) return P264B200_EINVAL;
    p264b200_gopdec *d = new (std::nothrow) p264b200_gopdec;
    if (!d) return P264B200_ENOMEM;
    std::vector<uint8_t> params;
    scan(annexb, bytes, d->gops, params);
    if (d->gops.empty()) {
        fprintf(stderr, "p264b200_gopdec_open: the stream has no IDR picture\n");
        delete d;
        return P264B200_EBITSTREAM;
    }
    if ((size_t)lanes > d->gops.size()) lanes = (int)d->gops.size();
    d->lanes = lanes;
    d->lane_stream.resize(lanes);
    d->fifo.resize(lanes);
    d->produced.resize(lanes);
    for (int l = 0; l < lanes; l++) d->lane_stream[l] = params;
    for (size_t gi = 0; gi < d->gops.size(); gi++) {
        std::vector<uint8_t> &s = d->lane_stream[gi % lanes];
        s.insert(s.end(), annexb + d->gops[gi].begin, annexb + d->gops[gi].end);
    }
    p264b200_multi_cfg cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.device = device;
    cfg.n_streams = lanes;
    cfg.n_threads = threads;
    int r = p264b200_multi_open(&d->multi, &cfg);
    for (int l = 0; l < lanes && r >= 0; l++) {
        d->lane_stream[l].resize(d->lane_stream[l].size() + 16, 0);   // scan slack
        r = p264b200_multi_set_stream(d->multi, l, d->lane_stream[l].data(), d->lane_stream[l].size() - 16);
    }
    if (r < 0) {
        p264b200_gopdec_close(d);
        return r;
    }
    *out = d;
    return lanes;
}

void p264b200_gopdec_close(p264b200_gopdec *d)
{
    if (!d) return;
    if (d->multi) p264b200_multi_close(d->multi);
    delete d;
}

int p264b200_gopdec_gops(const p264b200_gopdec *d) { return d ? (int)d->gops.size() : 0; }

int p264b200_gopdec_next(p264b200_gopdec *d, const uint8_t **picture, int *width, int *height)
{
    if (!d || !picture) return P264B200_EINVAL;
    for (;;) {
        while (d->gop_i < d->gops.size() && d->pic_i >= d->gops[d->gop_i].pictures) d->gop_i++, d->pic_i = 0;
        if (d->gop_i >= d->gops.size()) return 0;
        std::deque<std::vector<uint8_t>> &q = d->fifo[d->gop_i % d->lanes];
        // a lane delivers its GOPs in order, so the front of its queue is the picture the cursor points at
        if (!q.empty()) {
            d->current.swap(q.front());
            q.pop_front();
            d->pic_i++;
            *picture = d->current.data();
            if (width) *width = d->width;
            if (height) *height = d->height;
            return 1;
        }
        if (d->drained) return 0;   // the stream ended inside a GOP (truncated file): nothing more to deliver
        const int r = p264b200_multi_step(d->multi, d->produced.data());
        if (r < 0) return r;
        if (r == 0) {
            d->drained = true;
            continue;
        }
        for (int l = 0; l < d->lanes; l++) {
            if (!d->produced[l]) continue;
            const uint8_t *pic = p264b200_multi_picture(d->multi, l, &d->width, &d->height);
            if (pic) d->fifo[l].emplace_back(pic, pic + (size_t)d->width * d->height * 3 / 2);
        }
    }
}

}  // extern "C"
