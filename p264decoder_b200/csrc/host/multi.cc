// Multi-stream decoder: N independent Annex-B byte streams, one engine lane each, one batched GPU
// reconstruction per step (the host syntax front-end as a multi-threaded
// fast path and a multi-stream driver).  The reference has no equivalent -- its CLI decodes one
// stream on one core (p264decoder.c:164-381); running N copies of it is the CPU baseline.
//
// Per step every stream that still has data parses NAL units up to its next complete picture
// (worker threads, one Parser per stream: decoder/*.c is re-entrant per handle and so is this);
// the FrameSyntax of all streams goes down in ONE p264b200_stage_frames call, one
// p264b200_recon_step reconstructs all lanes, one p264b200_frames_download brings the pictures back
// as tight I420 images in pinned memory.  Streams that have ended keep their lane but are staged as
// an idle picture (no P slice, no intra macroblocks, no deblocking: every kernel skips the lane).
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "../../../include/p264b200_host.h"
#include "parser.h"

using namespace p264b200;

namespace {

struct Stream {
    Parser *parser = nullptr;
    const uint8_t *data = nullptr;
    size_t bytes = 0, pos = 0;
    std::vector<uint8_t> payload;   // unescaped NAL
    p264b200_frame_syntax fs;       // last parsed picture (parser-owned buffers)
    bool have_fs = false;           // fs was ever filled
    bool produced = false;          // ... in the current step
    bool ended = true;
    int err = 0;
    long pictures = 0;
    double t_scan = 0, t_nal = 0;   // profiling: seconds in the Annex-B scan + unescape / in Parser::nal
};

}  // namespace

struct p264b200_multi {
    p264b200_multi_cfg cfg;
    std::vector<Stream> streams;
    p264b200_engine *engine = nullptr;
    int mb_w = 0, mb_h = 0, ring = 0;
    uint8_t *out[2] = {nullptr, nullptr};   // pinned: [n_streams][picture_bytes], one per pipeline stage
    std::vector<uint8_t> delivered[2];      // which streams have a picture in out[i]
    int cur = 0, shown = -1, par = 0;       // stage in flight / stage the caller may read / stage being filled
    bool inflight = false, drained = false;
    size_t picture_bytes = 0;
    std::vector<p264b200_frame_syntax> batch;
    std::vector<int32_t> slots;
    // worker pool
    std::vector<std::thread> workers;
    std::mutex mu;
    std::condition_variable cv_go, cv_done;
    long generation = 0;
    int running = 0;
    bool quit = false;
    std::atomic<int> next_stream{0};
    void (*job)(p264b200_multi *, int) = nullptr;
    // pinned staging arenas laid out like the engine's staging area: [stream][n_mb] records, [stream][coef_cap] levels
    p264b200_mb *arena_mbs[2] = {nullptr, nullptr};
    int16_t *arena_coefs[2] = {nullptr, nullptr};
    size_t n_mb = 0, coef_cap = 0;
    // P264B200_MULTI_PROF=1: wall time per phase, printed at close
    bool prof = false;
    double t_parse = 0, t_stage = 0, t_recon = 0, t_down = 0;
    long steps = 0;
    bool first_done = false;
};

namespace {

double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// parse stream s up to its next complete picture
void parse_one(Stream &st)
{
    st.produced = false;
    if (st.ended || st.err) return;
    size_t start, n;
    for (;;) {
        const double ta = now_s();
        if (!p264b200_annexb_next(st.data, st.bytes, &st.pos, &start, &n)) break;
        int type = 0, ref_idc = 0;
        const int len = nal_unescape(st.data + start, (int)n, st.payload.data(), &type, &ref_idc);
        const double tb = now_s();
        st.t_scan += tb - ta;
        if (len < 0) continue;  // the reference CLI ignores undecodable NAL units as well
        int got = 0;
        p264b200_frame_syntax fs;
        const int r = st.parser->nal(type, ref_idc, st.payload.data(), len, &fs, &got);
        if (st.pictures) st.t_nal += now_s() - tb;
        if (r < 0) {
            st.err = r;
            return;
        }
        if (got) {
            st.fs = fs;
            st.have_fs = st.produced = true;
            st.pictures++;
            return;
        }
    }
    st.ended = true;
}

void worker_loop(p264b200_multi *m)
{
    long seen = 0;
    for (;;) {
        {
            std::unique_lock<std::mutex> lk(m->mu);
            m->cv_go.wait(lk, [&] { return m->quit || m->generation != seen; });
            if (m->quit) return;
            seen = m->generation;
        }
        for (;;) {
            const int s = m->next_stream.fetch_add(1);
            if (s >= (int)m->streams.size()) break;
            m->job(m, s);
        }
        {
            std::lock_guard<std::mutex> lk(m->mu);
            if (--m->running == 0) m->cv_done.notify_one();
        }
    }
}

void job_parse(p264b200_multi *m, int s) { parse_one(m->streams[s]); }

// the parser's buffers are ordinary memory; the picture moves into the pinned arena slot of its lane
void job_copy(p264b200_multi *m, int s)
{
    Stream &st = m->streams[s];
    if (!st.produced) return;
    p264b200_mb *mbs = m->arena_mbs[m->par] + (size_t)s * m->n_mb;
    int16_t *coefs = m->arena_coefs[m->par] + (size_t)s * m->coef_cap;
    memcpy(mbs, st.fs.mbs, m->n_mb * sizeof(p264b200_mb));
    if (st.fs.hdr.n_coef) memcpy(coefs, st.fs.coefs, (size_t)st.fs.hdr.n_coef * sizeof(int16_t));
    st.fs.mbs = mbs;
    st.fs.coefs = coefs;
}

void run_all(p264b200_multi *m, void (*job)(p264b200_multi *, int))
{
    if (m->workers.empty()) {
        for (int s = 0; s < (int)m->streams.size(); s++) job(m, s);
        return;
    }
    {
        std::lock_guard<std::mutex> lk(m->mu);
        m->job = job;
        m->next_stream.store(0);
        m->running = (int)m->workers.size();
        m->generation++;
    }
    m->cv_go.notify_all();
    std::unique_lock<std::mutex> lk(m->mu);
    m->cv_done.wait(lk, [&] { return m->running == 0; });
}

int ensure_engine(p264b200_multi *m)
{
    int mb_w = 0, mb_h = 0, ring = 0;
    for (auto &st : m->streams) {
        if (!st.produced) continue;
        const int w = st.parser->mb_w(), h = st.parser->mb_h();
        if (mb_w && (w != mb_w || h != mb_h)) {
            fprintf(stderr, "p264b200_multi: all streams of one batch must have the same coded size (%dx%d vs %dx%d macroblocks)\n", w, h,
                    mb_w, mb_h);
            return P264B200_EINVAL;
        }
        mb_w = w, mb_h = h;
        if (st.parser->ring_size() > ring) ring = st.parser->ring_size();
    }
    if (!mb_w) return 0;
    if (m->engine) {
        if (mb_w != m->mb_w || mb_h != m->mb_h || ring > m->ring) {
            fprintf(stderr, "p264b200_multi: coded size / DPB size changed mid-stream; not supported by the batched decoder\n");
            return P264B200_EINVAL;
        }
        return 0;
    }
    p264b200_engine_cfg cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.device = m->cfg.device;
    cfg.lanes = (int)m->streams.size();
    cfg.mb_w = mb_w, cfg.mb_h = mb_h;
    cfg.n_slots = ring;
    cfg.stage_steps = 2;
    const double tc0 = now_s();
    const int r = p264b200_engine_create(&m->engine, &cfg);
    if (m->prof) fprintf(stderr, "p264b200_multi: engine creation (incl. CUDA context) %.3f s\n", now_s() - tc0);
    if (r < 0) {
        fprintf(stderr, "p264b200_multi: GPU engine creation failed (%d): %s\n", r, p264b200_last_error());
        return r;
    }
    m->mb_w = mb_w, m->mb_h = mb_h, m->ring = ring;
    m->picture_bytes = (size_t)(16 * mb_w) * (16 * mb_h) * 3 / 2;
    m->n_mb = (size_t)mb_w * mb_h;
    m->coef_cap = (m->n_mb * 408 + 7) & ~(size_t)7;   // the engine's dense worst case (coef_capacity = 0)
    for (int i = 0; i < 2; i++) {
        m->out[i] = (uint8_t *)p264b200_host_alloc(m->picture_bytes * m->streams.size());
        m->arena_mbs[i] = (p264b200_mb *)p264b200_host_alloc(m->streams.size() * m->n_mb * sizeof(p264b200_mb));
        m->arena_coefs[i] = (int16_t *)p264b200_host_alloc(m->streams.size() * m->coef_cap * sizeof(int16_t));
        if (!m->out[i] || !m->arena_mbs[i] || !m->arena_coefs[i]) return P264B200_ENOMEM;
    }
    if (m->prof) fprintf(stderr, "p264b200_multi: ... + pinned arenas %.3f s\n", now_s() - tc0);
    return 0;
}

}  // namespace

extern "C" {

int p264b200_multi_open(p264b200_multi **out, const p264b200_multi_cfg *cfg)
{
    if (!out || !cfg || cfg->n_streams < 1 || cfg->n_streams > 256) return P264B200_EINVAL;
    if (p264b200_device_count() <= 0) {
        fprintf(stderr, "p264b200_multi_open: no CUDA device available and this build has no CPU reconstruction path\n");
        return P264B200_ENODEV;
    }
    p264b200_multi *m = new (std::nothrow) p264b200_multi;
    if (!m) return P264B200_ENOMEM;
    m->cfg = *cfg;
    m->streams.resize(cfg->n_streams);
    for (auto &st : m->streams) {
        st.parser = new (std::nothrow) Parser();
        if (!st.parser) {
            p264b200_multi_close(m);
            return P264B200_ENOMEM;
        }
        st.parser->verbose = 0;
    }
    m->batch.resize(cfg->n_streams);
    m->delivered[0].assign(cfg->n_streams, 0);
    m->delivered[1].assign(cfg->n_streams, 0);
    m->slots.resize(cfg->n_streams);
    m->prof = getenv("P264B200_MULTI_PROF") != nullptr;
    int nt = cfg->n_threads > 0 ? cfg->n_threads : (int)std::thread::hardware_concurrency();
    if (nt > cfg->n_streams) nt = cfg->n_streams;
    if (nt > 1)
        for (int i = 0; i < nt; i++) m->workers.emplace_back(worker_loop, m);
    *out = m;
    return P264B200_OK;
}

void p264b200_multi_close(p264b200_multi *m)
{
    if (!m) return;
    {
        std::lock_guard<std::mutex> lk(m->mu);
        m->quit = true;
    }
    m->cv_go.notify_all();
    for (auto &t : m->workers) t.join();
    if (m->prof && m->steps) {
        double ts = 0, tn = 0;
        long np = 0;
        for (auto &st : m->streams) ts += st.t_scan, tn += st.t_nal, np += st.pictures;
        fprintf(stderr, "p264b200_multi: per picture: Annex-B scan + unescape %.3f ms, Parser::nal %.3f ms\n", 1e3 * ts / (np ? np : 1), 1e3 * tn / (np ? np : 1));
    }
    if (m->prof && m->steps)
        fprintf(stderr, "p264b200_multi: %ld steps; per step: parse + copy %.3f ms (overlaps the GPU), submit %.3f ms, wait for the GPU %.3f ms\n", m->steps,
                1e3 * m->t_parse / m->steps, 1e3 * m->t_stage / m->steps, 1e3 * m->t_down / m->steps);
    if (m->engine) p264b200_engine_destroy(m->engine);
    for (auto &st : m->streams) delete st.parser;
    for (int i = 0; i < 2; i++) {
        p264b200_host_free(m->out[i]);
        p264b200_host_free(m->arena_mbs[i]);
        p264b200_host_free(m->arena_coefs[i]);
    }
    delete m;
}

int p264b200_multi_set_stream(p264b200_multi *m, int s, const uint8_t *annexb, size_t bytes)
{
    if (!m || s < 0 || s >= (int)m->streams.size() || !annexb) return P264B200_EINVAL;
    Stream &st = m->streams[s];
    st.data = annexb, st.bytes = bytes, st.pos = 0;
    st.payload.resize(bytes + 16);
    st.ended = bytes == 0;
    return P264B200_OK;
}

// parse the next picture of every stream into arena `par`; returns pictures parsed (0 = all streams ended) or < 0
static int parse_step(p264b200_multi *m, int par)
{
    const double t0 = now_s();
    run_all(m, job_parse);
    int n_pic = 0;
    for (auto &st : m->streams) {
        if (st.err) return st.err;
        n_pic += st.produced;
    }
    if (n_pic) {
        const int r = ensure_engine(m);
        if (r < 0) return r;
        m->par = par;
        run_all(m, job_copy);
    }
    if (m->first_done) m->t_parse += now_s() - t0;
    return n_pic;
}

// hand the parsed step (arena `par`) to the GPU: stage + reconstruct + download, all asynchronous
static int submit_step(p264b200_multi *m, int par)
{
    const double t0 = now_s();
    const int n = (int)m->streams.size();
    const p264b200_frame_syntax *any = nullptr;
    for (auto &st : m->streams)
        if (st.produced) any = &st.fs;
    for (int s = 0; s < n; s++) {
        Stream &st = m->streams[s];
        m->delivered[par][s] = st.produced;
        if (st.produced) {
            m->batch[s] = st.fs;
            m->slots[s] = st.fs.hdr.dst_slot;
        } else {
            // idle lane: nothing to reconstruct; every kernel skips it (the border pass rewrites the same samples)
            p264b200_frame_syntax idle = st.have_fs ? st.fs : *any;
            idle.mbs = m->arena_mbs[par] + (size_t)s * m->n_mb;   // (stale or never written: nothing reads the records of an idle lane)
            idle.coefs = m->arena_coefs[par] + (size_t)s * m->coef_cap;
            idle.hdr.slice_type = P264B200_SLICE_I;
            idle.hdr.n_intra = 0;
            idle.hdr.deblock = 0;
            idle.hdr.n_coef = 0;
            idle.hdr.num_ref = 0;
            if (!st.have_fs) idle.hdr.dst_slot = 0;
            m->batch[s] = idle;
            m->slots[s] = idle.hdr.dst_slot;
        }
    }
    int r;
    if ((r = p264b200_stage_frames(m->engine, par, n, m->batch.data())) < 0 || (r = p264b200_recon_step(m->engine, par, n)) < 0 ||
        (r = p264b200_frames_download(m->engine, n, m->slots.data(), m->out[par], m->picture_bytes)) < 0) {
        fprintf(stderr, "p264b200_multi: GPU reconstruction failed (%d): %s\n", r, p264b200_last_error());
        return r;
    }
    if (m->first_done) m->t_stage += now_s() - t0;
    return 0;
}

// Two-deep pipeline: while the GPU reconstructs step i (arena / output buffer i & 1), the host threads parse step
// i + 1 into the other arena; the call then waits for step i, submits step i + 1 and delivers the pictures of step i.
int p264b200_multi_step(p264b200_multi *m, uint8_t *produced)
{
    if (!m) return P264B200_EINVAL;
    int r;
    if (!m->inflight) {
        if (m->drained) return 0;
        if ((r = parse_step(m, m->cur)) <= 0) {
            m->drained = r == 0;
            return r;
        }
        if ((r = submit_step(m, m->cur)) < 0) return r;
        m->inflight = true;
    }
    const int cur = m->cur, nxt = cur ^ 1;
    const int n_next = parse_step(m, nxt);   // overlaps the GPU work of step `cur`
    if (n_next < 0) return n_next;
    const double t0 = now_s();
    if ((r = p264b200_engine_sync(m->engine)) < 0) {
        fprintf(stderr, "p264b200_multi: GPU reconstruction failed (%d): %s\n", r, p264b200_last_error());
        return r;
    }
    if (m->first_done) m->t_down += now_s() - t0, m->steps++;
    m->first_done = true;
    int n_pic = 0;
    for (size_t s = 0; s < m->streams.size(); s++) n_pic += m->delivered[cur][s];
    if (produced) memcpy(produced, m->delivered[cur].data(), m->streams.size());
    m->shown = cur;
    if (n_next > 0) {
        if ((r = submit_step(m, nxt)) < 0) return r;
        m->cur = nxt;
    } else {
        m->inflight = false;
        m->drained = true;
    }
    return n_pic;
}

const uint8_t *p264b200_multi_picture(const p264b200_multi *m, int s, int *width, int *height)
{
    if (!m || s < 0 || s >= (int)m->streams.size() || m->shown < 0 || !m->delivered[m->shown][s]) return nullptr;
    if (width) *width = 16 * m->mb_w;
    if (height) *height = 16 * m->mb_h;
    return m->out[m->shown] + (size_t)s * m->picture_bytes;
}

}  // extern "C"
