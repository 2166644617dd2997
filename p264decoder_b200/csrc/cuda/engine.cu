// Host side of the frame-level C-ABI (include/p264b200_recon.h): device frame ring, syntax
// staging and the per-picture kernel sequence that replaces p264_slice_decode steps [3]-[4]
// (decoder/decoder.c:623-661).  There is deliberately no CPU path in this file: without a
// usable CUDA device every entry point fails with P264B200_ENODEV.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <algorithm>
#include <utility>
#include <vector>

#define P264B200_DEFINE_KERNELS 1
#include "border.cuh"
#include "common.cuh"
#include "deblock.cuh"
#include "expand_v2.cuh"
#include "md5.cuh"
#include "recon_inter.cuh"
#include "recon_intra.cuh"

using namespace p264b200;

namespace {
thread_local char g_err[512] = "";
void set_err(const char *what, cudaError_t e)
{
    snprintf(g_err, sizeof(g_err), "%s: %s", what, e == cudaSuccess ? "" : cudaGetErrorString(e));
}
#define CK(call)                                  \
    do {                                          \
        cudaError_t e_ = (call);                  \
        if (e_ != cudaSuccess) {                  \
            set_err(#call, e_);                   \
            return P264B200_ECUDA;                \
        }                                         \
    } while (0)

enum { K_INTER = 0, K_INTRA, K_DEBLOCK, K_BORDER, K_DEBLOCK_BS, K_COUNT };
}  // namespace

struct p264b200_engine {
    p264b200_engine_cfg cfg;
    Geometry g;
    cudaStream_t stream = nullptr;                 // compute (kernels)
    // Host-buffer path: syntax uploads, kernels and picture downloads run on three streams so that
    // H2D of step n+1, reconstruction of step n and D2H of step n-1 overlap (PCIe is full duplex).
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
    // the boundary-strength pre-pass reads syntax only, so it runs beside recon_inter on its own stream and joins before deblock
    cudaStream_t s_side = nullptr;
    cudaEvent_t ev_side_fork = nullptr, ev_side_join = nullptr;
    std::vector<cudaEvent_t> ev_staged;            // [step] last upload into that staging slot
    std::vector<cudaEvent_t> ev_recon;             // [step] reconstruction that consumed that slot
    std::vector<cudaEvent_t> ev_d2h_slot;          // [frame slot] last download of that ring slot (any lane)
    cudaEvent_t ev_compute = nullptr;              // last reconstruction issued
    std::vector<uint8_t> slot_dst;                 // [step][lane] ring slot that picture is reconstructed into
    std::vector<uint8_t> desc_queued;              // [step][lane] an upload from the pinned h_descs entry may still be queued on s_h2d
    bool h2d_busy = false, d2h_busy = false;
    // frame store
    uint8_t *d_y = nullptr, *d_c = nullptr;
    // staging: [step][lane]
    p264b200_mb *d_mbs = nullptr;
    int16_t *d_coefs = nullptr;
    FrameDesc *d_descs = nullptr, *h_descs = nullptr;
    size_t coef_cap = 0;  // int16 per lane per step
    // FrameSyntax v2: packed pictures [step][lanes x blob capacity] + their section tables (lazy, first p264b200_stage_frames_v2)
    uint8_t *d_blob = nullptr;
    size_t blob_cap = 0;           // bytes per lane per step
    V2Desc *d_v2 = nullptr, *h_v2 = nullptr;
    uint8_t *d_out = nullptr;      // [lanes] tight I420 pictures for the batched download (lazy)
    size_t out_bytes = 0;
    PackSrc *d_pack = nullptr;     // [lanes][n_slots] plane origins
    uint8_t *d_md5 = nullptr;      // [lanes][16] digests (lazy)
    DeblockSide *d_bs = nullptr;   // [lanes][n_mb]
    int *d_intra = nullptr;        // [lanes][1 + 2*n_mb]: intra runs + per-MB done epochs
    int intra_epoch = 0;
    std::vector<int> slot_nintra;  // [step][lane]: intra macroblocks of the staged picture
    int *d_sync = nullptr;  // [groups][kSyncHdr] tickets + per-SM arrival counters, then [lanes][3*mb_h]: intra, luma deblock, chroma deblock wavefronts
    size_t sync_bytes = 0;
    std::vector<uint8_t> slot_flags;  // [step][lane]: bit0 intra MBs present, bit1 deblock on, bit2 P slice
    // timing
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool profile = false;
    std::vector<cudaEvent_t> prof_ev;   // pairs
    std::vector<int> prof_kind;
    size_t prof_used = 0;
    float prof_ms[K_COUNT] = {0, 0, 0, 0, 0};
    uint64_t prof_n[K_COUNT] = {0, 0, 0, 0, 0};
    uint64_t launches = 0;
    // lane groups: independent lanes are split over `n_groups` CUDA streams so that one group's
    // latency-bound wavefront kernels overlap with another group's MC kernel on the same SMs
    static constexpr int kMaxGroups = 8;
    static constexpr int kSyncHdr = 4 + kDbfSmSlots;  // per lane group: intra ticket, luma / chroma deblock tickets, pad, per-SM arrival counters
    int n_groups = 1;
    cudaStream_t gstream[kMaxGroups] = {};   // low priority: recon_inter of the group
    cudaStream_t hstream[kMaxGroups] = {};   // high priority: bS pre-pass, intra, deblock, border of the group
    cudaEvent_t ev_fork = nullptr, ev_join[kMaxGroups] = {}, ev_mc[kMaxGroups] = {}, ev_hi_done[kMaxGroups] = {};
    int dbf_pad_bytes = 0;   // P264B200_DBF_PAD (KB): dynamic shared memory added to deblock CTAs = fewer of them per SM, room for recon_inter CTAs
    bool groups_dirty = false;  // group streams hold work the main stream has not joined yet
    int inter_variant = 0;  // P264B200_INTER_VARIANT: CTA shape of recon_inter.  0 (default) = 8x8-macroblock tiles, 384 threads x 3 CTAs per SM
                            // (56 registers): 1.47 ms at 256 lanes; 1 = 512 threads x 2 (64 registers): 1.75 ms; 2 = 8x16 tiles, 576 x 2: 1.74 ms;
                            // 3 = 8x4 tiles, 192 x 6: 1.53 ms.  (4 CTAs per SM at 40 / 48 registers: 2.2 ms.)
    int dbf_variant = 0;   // P264B200_DBF_VARIANT: rows per deblock CTA x CTAs per SM.  0 (default) = 8 x 2 (91 registers, 16 filtering warps per SM): 1.150 ms
                           // at 256 lanes; 1 = 6 x 3 (80 registers, 18 warps): 1.153; 2 = 7 x 3 (72 registers, 21 warps): 1.173; also measured 4 x 4: 1.228,
                           // 10 x 2: 1.271 -- more resident chains do not help, the kernel is bound by the ALU / L1 pipes its chains share
    bool no_side = true;   // P264B200_NO_SIDE=0: run the boundary-strength pre-pass on a side stream beside recon_inter (measured: step 2.780 ->
                           // 2.764 ms, but the pre-pass then shares the SMs for the whole 1.5 ms and the per-kernel profile stops adding up; off by default)
    int dbg = 0;  // P264B200_DBG: timing experiments only (results are wrong when set)
    int trace_ticket = -1;  // P264B200_TRACE: deblock CTA (by ticket) whose per-step cycle marks are recorded

    uint8_t *plane(int lane, int slot, int c) const
    {
        const size_t f = (size_t)lane * cfg.n_slots + slot;
        if (c == 0) return d_y + f * g.y_plane + (size_t)kLumaPad * g.y_stride + kLumaPad;
        return d_c + (f * 2 + (c - 1)) * g.c_plane + (size_t)kChromaPad * g.c_stride + kChromaPad;
    }
};

namespace {

struct ProfScope {
    p264b200_engine *e;
    int kind;
    size_t idx = 0;
    bool on;
    cudaStream_t st;
    ProfScope(p264b200_engine *e_, int k, cudaStream_t s = nullptr) : e(e_), kind(k), on(e_->profile), st(s ? s : e_->stream)
    {
        e->launches++;
        if (!on) return;
        if (e->prof_used + 2 > e->prof_ev.size()) {
            for (int i = 0; i < 2; i++) {
                cudaEvent_t ev;
                cudaEventCreate(&ev);
                e->prof_ev.push_back(ev);
            }
            e->prof_kind.push_back(kind);
        }
        idx = e->prof_used;
        e->prof_kind[idx / 2] = kind;
        e->prof_used += 2;
        cudaEventRecord(e->prof_ev[idx], st);
    }
    ~ProfScope()
    {
        if (on) cudaEventRecord(e->prof_ev[idx + 1], st);
    }
};

// make the main stream wait for everything the lane-group streams were given
cudaError_t join_groups(p264b200_engine *e)
{
    if (!e->groups_dirty) return cudaSuccess;
    for (int gi = 0; gi < e->n_groups; gi++) {
        cudaError_t err = cudaEventRecord(e->ev_join[gi], e->hstream[gi]);
        if (err == cudaSuccess) err = cudaStreamWaitEvent(e->stream, e->ev_join[gi], 0);
        if (err != cudaSuccess) return err;
    }
    e->groups_dirty = false;
    return cudaSuccess;
}

void prof_collect(p264b200_engine *e)
{
    for (size_t i = 0; i + 1 < e->prof_used; i += 2) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, e->prof_ev[i], e->prof_ev[i + 1]) == cudaSuccess) {
            e->prof_ms[e->prof_kind[i / 2]] += ms;
            e->prof_n[e->prof_kind[i / 2]]++;
        }
    }
    e->prof_used = 0;
}

}  // namespace

extern "C" {

const char *p264b200_last_error(void) { return g_err; }
int p264b200_abi_version(void) { return P264B200_ABI_VERSION; }

int p264b200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

void *p264b200_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
void p264b200_host_free(void *p)
{
    if (p) cudaFreeHost(p);
}

void p264b200_engine_destroy(p264b200_engine *e)
{
    if (!e) return;
    cudaSetDevice(e->cfg.device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    cudaFree(e->d_y);
    cudaFree(e->d_c);
    cudaFree(e->d_mbs);
    cudaFree(e->d_coefs);
    cudaFree(e->d_descs);
    cudaFree(e->d_sync);
    cudaFree(e->d_bs);
    cudaFree(e->d_out);
    cudaFree(e->d_pack);
    cudaFree(e->d_md5);
    cudaFree(e->d_intra);
    cudaFree(e->d_blob);
    cudaFree(e->d_v2);
    if (e->h_v2) cudaFreeHost(e->h_v2);
    if (e->h_descs) cudaFreeHost(e->h_descs);
    for (auto ev : e->prof_ev) cudaEventDestroy(ev);
    if (e->s_h2d) cudaStreamSynchronize(e->s_h2d);
    if (e->s_d2h) cudaStreamSynchronize(e->s_d2h);
    for (auto *vec : {&e->ev_staged, &e->ev_recon, &e->ev_d2h_slot})
        for (auto ev : *vec)
            if (ev) cudaEventDestroy(ev);
    if (e->ev_compute) cudaEventDestroy(e->ev_compute);
    if (e->s_side) cudaStreamSynchronize(e->s_side), cudaStreamDestroy(e->s_side);
    if (e->ev_side_fork) cudaEventDestroy(e->ev_side_fork);
    if (e->ev_side_join) cudaEventDestroy(e->ev_side_join);
    if (e->s_h2d) cudaStreamDestroy(e->s_h2d);
    if (e->s_d2h) cudaStreamDestroy(e->s_d2h);
    if (e->ev0) cudaEventDestroy(e->ev0);
    if (e->ev1) cudaEventDestroy(e->ev1);
    if (e->ev_fork) cudaEventDestroy(e->ev_fork);
    for (int gi = 0; gi < p264b200_engine::kMaxGroups; gi++) {
        if (e->ev_join[gi]) cudaEventDestroy(e->ev_join[gi]);
        if (e->ev_mc[gi]) cudaEventDestroy(e->ev_mc[gi]);
        if (e->ev_hi_done[gi]) cudaEventDestroy(e->ev_hi_done[gi]);
        if (e->gstream[gi]) cudaStreamDestroy(e->gstream[gi]);
        if (e->hstream[gi]) cudaStreamDestroy(e->hstream[gi]);
    }
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
}

int p264b200_engine_create(p264b200_engine **out, const p264b200_engine_cfg *cfg)
{
    if (!out || !cfg || cfg->lanes < 1 || cfg->mb_w < 1 || cfg->mb_h < 1 || cfg->n_slots < 1 || cfg->n_slots > 17 ||
        cfg->stage_steps < 1) {
        set_err("p264b200_engine_create: bad configuration", cudaSuccess);
        return P264B200_EINVAL;
    }
    if (p264b200_device_count() <= cfg->device) {
        set_err("p264b200_engine_create: no CUDA device (this engine has no CPU fallback)", cudaSuccess);
        return P264B200_ENODEV;
    }
    CK(cudaSetDevice(cfg->device));
    p264b200_engine *e = new (std::nothrow) p264b200_engine;
    if (!e) return P264B200_ENOMEM;
    e->cfg = *cfg;
    if (const char *d = getenv("P264B200_DBG")) e->dbg = atoi(d);
    if (const char *d = getenv("P264B200_TRACE")) e->trace_ticket = atoi(d);
    if (const char *d = getenv("P264B200_NO_SIDE")) e->no_side = atoi(d) != 0;
    if (const char *d = getenv("P264B200_DBF_VARIANT")) e->dbf_variant = atoi(d);
    // measured (256 lanes x 1080p): 1 group 3.05 ms per step; 2 groups pipelined across steps (one group's deblock beside the
    // other's recon_inter, priority streams) 3.15 ms; with deblock held to one CTA per SM (P264B200_DBF_PAD=100) 3.46 ms; 4 groups
    // 3.63 ms -- both kernels lean on the same L1 / shared-memory data pipe (82 % and 67 % alone), so sharing the SMs buys nothing
    e->n_groups = 1;
    if (const char *gq = getenv("P264B200_GROUPS")) e->n_groups = atoi(gq);
    if (e->n_groups < 1) e->n_groups = 1;
    if (e->n_groups > p264b200_engine::kMaxGroups) e->n_groups = p264b200_engine::kMaxGroups;
    if (e->n_groups > cfg->lanes) e->n_groups = cfg->lanes;
    Geometry &g = e->g;
    g.mb_w = cfg->mb_w;
    g.mb_h = cfg->mb_h;
    g.width = 16 * cfg->mb_w;
    g.height = 16 * cfg->mb_h;
    g.y_stride = (g.width + 2 * kLumaPad + 127) & ~127;
    g.c_stride = g.y_stride / 2;
    g.y_rows = g.height + 2 * kLumaPad;
    g.c_rows = g.height / 2 + 2 * kChromaPad;
    g.y_plane = (size_t)g.y_stride * g.y_rows;
    g.c_plane = (size_t)g.c_stride * g.c_rows;
    const size_t n_mb = (size_t)g.mb_w * g.mb_h;
    e->coef_cap = cfg->coef_capacity ? cfg->coef_capacity : n_mb * 408;
    e->coef_cap = (e->coef_cap + 7) & ~(size_t)7;
    const size_t frames = (size_t)cfg->lanes * cfg->n_slots;
    const size_t slots = (size_t)cfg->stage_steps * cfg->lanes;
    e->sync_bytes = ((size_t)p264b200_engine::kSyncHdr * p264b200_engine::kMaxGroups + (size_t)cfg->lanes * 3 * g.mb_h) * sizeof(int);
    int rc = P264B200_OK;
    auto fail = [&](const char *what, cudaError_t err) {
        set_err(what, err);
        rc = err == cudaErrorMemoryAllocation ? P264B200_ENOMEM : P264B200_ECUDA;
    };
    cudaError_t err;
    if ((err = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking)) != cudaSuccess) fail("stream", err);
    if (!rc && (err = cudaMalloc(&e->d_y, frames * g.y_plane)) != cudaSuccess) fail("cudaMalloc luma store", err);
    if (!rc && (err = cudaMalloc(&e->d_c, frames * 2 * g.c_plane)) != cudaSuccess) fail("cudaMalloc chroma store", err);
    if (!rc && (err = cudaMalloc(&e->d_mbs, slots * n_mb * sizeof(p264b200_mb))) != cudaSuccess) fail("cudaMalloc mbs", err);
    if (!rc && (err = cudaMalloc(&e->d_coefs, slots * e->coef_cap * sizeof(int16_t))) != cudaSuccess) fail("cudaMalloc coefs", err);
    if (!rc && (err = cudaMalloc(&e->d_descs, slots * sizeof(FrameDesc))) != cudaSuccess) fail("cudaMalloc descs", err);
    if (!rc && (err = cudaMallocHost(&e->h_descs, slots * sizeof(FrameDesc))) != cudaSuccess) fail("cudaMallocHost descs", err);
    if (!rc && (err = cudaMalloc(&e->d_sync, e->sync_bytes)) != cudaSuccess) fail("cudaMalloc sync", err);
    if (!rc && (err = cudaMalloc(&e->d_bs, (size_t)cfg->lanes * n_mb * sizeof(DeblockSide))) != cudaSuccess) fail("cudaMalloc bs", err);
    if (!rc && (err = cudaMalloc(&e->d_intra, (size_t)cfg->lanes * (1 + 2 * n_mb) * sizeof(int))) != cudaSuccess) fail("cudaMalloc intra work", err);
    if (!rc && (err = cudaMemset(e->d_intra, 0, (size_t)cfg->lanes * (1 + 2 * n_mb) * sizeof(int))) != cudaSuccess) fail("memset intra work", err);
    e->slot_nintra.assign(slots, 0);
    if (!rc && (err = cudaStreamCreateWithFlags(&e->s_h2d, cudaStreamNonBlocking)) != cudaSuccess) fail("stream", err);
    if (!rc && (err = cudaStreamCreateWithFlags(&e->s_d2h, cudaStreamNonBlocking)) != cudaSuccess) fail("stream", err);
    if (!rc && (err = cudaEventCreateWithFlags(&e->ev_compute, cudaEventDisableTiming)) != cudaSuccess) fail("event", err);
    if (!rc && (err = cudaStreamCreateWithFlags(&e->s_side, cudaStreamNonBlocking)) != cudaSuccess) fail("stream", err);
    if (!rc && (err = cudaEventCreateWithFlags(&e->ev_side_fork, cudaEventDisableTiming)) != cudaSuccess) fail("event", err);
    if (!rc && (err = cudaEventCreateWithFlags(&e->ev_side_join, cudaEventDisableTiming)) != cudaSuccess) fail("event", err);
    e->ev_staged.assign(cfg->stage_steps, nullptr);
    e->ev_recon.assign(cfg->stage_steps, nullptr);
    e->ev_d2h_slot.assign(cfg->n_slots, nullptr);
    for (auto *vec : {&e->ev_staged, &e->ev_recon, &e->ev_d2h_slot})
        for (auto &ev : *vec)
            if (!rc && (err = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) != cudaSuccess) fail("event", err);
    e->slot_dst.assign(slots, 0);
    e->desc_queued.assign(slots, 0);
    if (!rc && (err = cudaEventCreate(&e->ev0)) != cudaSuccess) fail("event", err);
    if (!rc && (err = cudaEventCreate(&e->ev1)) != cudaSuccess) fail("event", err);
    if (!rc && (err = cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming)) != cudaSuccess) fail("event", err);
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (const char *v = getenv("P264B200_INTER_VARIANT")) e->inter_variant = atoi(v);
    {
        const std::pair<const void *, int> kern[] = {
            {(const void *)recon_inter_kernel<384, 3, 8>, (int)sizeof(InterSmem<8>)}, {(const void *)recon_inter_kernel<512, 2, 8>, (int)sizeof(InterSmem<8>)},
            {(const void *)recon_inter_kernel<576, 2, 16>, (int)sizeof(InterSmem<16>)}, {(const void *)recon_inter_kernel<192, 6, 4>, (int)sizeof(InterSmem<4>)}};
        for (auto &k : kern)
            if (!rc && (err = cudaFuncSetAttribute(k.first, cudaFuncAttributeMaxDynamicSharedMemorySize, k.second)) != cudaSuccess)
                fail("cudaFuncSetAttribute(recon_inter)", err);
    }
    if (const char *pad = getenv("P264B200_DBF_PAD")) e->dbf_pad_bytes = atoi(pad) * 1024;
    if (!rc && e->dbf_pad_bytes > 0 &&
        (err = cudaFuncSetAttribute(deblock_kernel<kDbfRows, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, e->dbf_pad_bytes)) != cudaSuccess)
        fail("cudaFuncSetAttribute", err);
    for (int gi = 0; gi < e->n_groups && !rc && e->n_groups > 1; gi++) {
        if ((err = cudaStreamCreateWithPriority(&e->gstream[gi], cudaStreamNonBlocking, prio_lo)) != cudaSuccess) fail("group stream", err);
        if (!rc && (err = cudaStreamCreateWithPriority(&e->hstream[gi], cudaStreamNonBlocking, prio_hi)) != cudaSuccess) fail("group stream", err);
        if (!rc && (err = cudaEventCreateWithFlags(&e->ev_hi_done[gi], cudaEventDisableTiming)) != cudaSuccess) fail("event", err);
        if (!rc && (err = cudaEventCreateWithFlags(&e->ev_join[gi], cudaEventDisableTiming)) != cudaSuccess) fail("event", err);
        if (!rc && (err = cudaEventCreateWithFlags(&e->ev_mc[gi], cudaEventDisableTiming)) != cudaSuccess) fail("event", err);
    }
    if (!rc) {
        // grey frames so that never-written slots are deterministic
        if ((err = cudaMemsetAsync(e->d_y, 128, frames * g.y_plane, e->stream)) != cudaSuccess) fail("memset", err);
        if (!rc && (err = cudaMemsetAsync(e->d_c, 128, frames * 2 * g.c_plane, e->stream)) != cudaSuccess) fail("memset", err);
        if (!rc && (err = cudaStreamSynchronize(e->stream)) != cudaSuccess) fail("sync", err);
    }
    if (rc) {
        p264b200_engine_destroy(e);
        return rc;
    }
    memset(e->h_descs, 0, slots * sizeof(FrameDesc));
    e->slot_flags.assign(slots, 0);
    *out = e;
    return P264B200_OK;
}

int p264b200_engine_geometry(const p264b200_engine *e, int32_t *luma_stride, int32_t *chroma_stride, int32_t *width,
                             int32_t *height)
{
    if (!e) return P264B200_EINVAL;
    if (luma_stride) *luma_stride = e->g.y_stride;
    if (chroma_stride) *chroma_stride = e->g.c_stride;
    if (width) *width = e->g.width;
    if (height) *height = e->g.height;
    return P264B200_OK;
}

}  // extern "C" (reopened below)

namespace {
int ensure_pack_table(p264b200_engine *e)
{
    if (e->d_pack) return P264B200_OK;
    std::vector<PackSrc> tab((size_t)e->cfg.lanes * e->cfg.n_slots);
    for (int l = 0; l < e->cfg.lanes; l++)
        for (int sl = 0; sl < e->cfg.n_slots; sl++)
            for (int c = 0; c < 3; c++) tab[(size_t)l * e->cfg.n_slots + sl].plane[c] = e->plane(l, sl, c);
    CK(cudaMalloc(&e->d_pack, tab.size() * sizeof(PackSrc)));
    CK(cudaMemcpy(e->d_pack, tab.data(), tab.size() * sizeof(PackSrc), cudaMemcpyHostToDevice));
    return P264B200_OK;
}

// validation + FrameDesc of one lane's picture (no copies)
int prepare_desc(p264b200_engine *e, int step, int lane, const p264b200_frame_syntax *fs)
{
    const p264b200_frame_hdr &h = fs->hdr;
    const Geometry &g = e->g;
    if (h.mb_w != g.mb_w || h.mb_h != g.mb_h || h.dst_slot < 0 || h.dst_slot >= e->cfg.n_slots || h.num_ref < 0 ||
        h.num_ref > kMaxRefs || h.n_coef > e->coef_cap || !fs->mbs || (h.n_coef && !fs->coefs)) {
        set_err("p264b200_stage_frame: syntax does not fit the engine", cudaSuccess);
        return P264B200_EINVAL;
    }
    if (h.slice_type == P264B200_SLICE_P && h.num_ref < 1) return P264B200_EINVAL;
    for (int i = 0; i < h.num_ref; i++)
        if (h.ref_slot[i] < 0 || h.ref_slot[i] >= e->cfg.n_slots) return P264B200_EINVAL;
    const size_t n_mb = (size_t)g.mb_w * g.mb_h;
    const size_t s = (size_t)step * e->cfg.lanes + lane;
    if (e->desc_queued[s]) {
        // The pinned descriptor is the source of an asynchronous copy that may not have run yet (a caller may stage a
        // step again as soon as the previous call returned): wait for this step's uploads before rewriting it.  One
        // wait per step and generation -- it also covers every other lane's entry of the step.
        if (cudaEventSynchronize(e->ev_staged[step]) != cudaSuccess) {
            set_err("cudaEventSynchronize(ev_staged)", cudaGetLastError());
            return P264B200_ECUDA;
        }
        memset(&e->desc_queued[(size_t)step * e->cfg.lanes], 0, e->cfg.lanes);
    }
    FrameDesc &d = e->h_descs[s];
    memset(&d, 0, sizeof(d));
    d.mbs = e->d_mbs + s * n_mb;
    d.coefs = e->d_coefs + s * e->coef_cap;
    for (int c = 0; c < 3; c++) d.cur[c] = e->plane(lane, h.dst_slot, c);
    for (int i = 0; i < h.num_ref; i++)
        for (int c = 0; c < 3; c++) d.ref[i][c] = e->plane(lane, h.ref_slot[i], c);
    d.row_progress = e->d_sync + p264b200_engine::kSyncHdr * p264b200_engine::kMaxGroups + (size_t)lane * 3 * g.mb_h;
    d.dbf_bs = e->d_bs + (size_t)lane * n_mb;
    d.intra_work = e->d_intra + (size_t)lane * (1 + 2 * n_mb);
    d.slice_type = h.slice_type;
    d.deblock = h.deblock;
    d.alpha_off = h.alpha_c0_offset;
    d.beta_off = h.beta_offset;
    d.chroma_qp_off = h.chroma_qp_index_offset;
    d.n_intra = h.n_intra;
    d.num_ref = h.num_ref;
    e->slot_nintra[s] = h.n_intra;
    e->slot_flags[s] = (uint8_t)((h.n_intra > 0) | ((h.deblock != 0) << 1) | ((h.slice_type == P264B200_SLICE_P) << 2));
    e->slot_dst[s] = (uint8_t)h.dst_slot;
    return P264B200_OK;
}
}  // namespace

extern "C" {

int p264b200_stage_frame(p264b200_engine *e, int step, int lane, const p264b200_frame_syntax *fs)
{
    if (!e || !fs || step < 0 || step >= e->cfg.stage_steps || lane < 0 || lane >= e->cfg.lanes) return P264B200_EINVAL;
    const int r = prepare_desc(e, step, lane, fs);
    if (r) return r;
    CK(cudaSetDevice(e->cfg.device));
    const size_t n_mb = (size_t)e->g.mb_w * e->g.mb_h;
    const size_t s = (size_t)step * e->cfg.lanes + lane;
    // the staging slot may still be read by the reconstruction that last used it
    CK(cudaStreamWaitEvent(e->s_h2d, e->ev_recon[step], 0));
    CK(cudaMemcpyAsync(e->d_mbs + s * n_mb, fs->mbs, n_mb * sizeof(p264b200_mb), cudaMemcpyHostToDevice, e->s_h2d));
    if (fs->hdr.n_coef)
        CK(cudaMemcpyAsync(e->d_coefs + s * e->coef_cap, fs->coefs, (size_t)fs->hdr.n_coef * sizeof(int16_t), cudaMemcpyHostToDevice, e->s_h2d));
    CK(cudaMemcpyAsync(e->d_descs + s, &e->h_descs[s], sizeof(FrameDesc), cudaMemcpyHostToDevice, e->s_h2d));
    e->desc_queued[s] = 1;
    CK(cudaEventRecord(e->ev_staged[step], e->s_h2d));
    e->h2d_busy = true;
    return P264B200_OK;
}

int p264b200_stage_frames(p264b200_engine *e, int step, int n, const p264b200_frame_syntax *fs)
{
    if (!e || !fs || n < 1 || n > e->cfg.lanes || step < 0 || step >= e->cfg.stage_steps) return P264B200_EINVAL;
    const size_t n_mb = (size_t)e->g.mb_w * e->g.mb_h;
    // host buffers laid out like the device staging area ([lane][n_mb] records, [lane][coef_capacity] levels):
    // three copies for the whole step instead of three per lane (the per-call cost of ~200 small copies,
    // not PCIe, was the limit of the host-buffer path)
    bool contiguous = true;
    for (int l = 1; l < n && contiguous; l++)
        contiguous = fs[l].mbs == fs[0].mbs + (size_t)l * n_mb && fs[l].coefs == fs[0].coefs + (size_t)l * e->coef_cap;
    if (!contiguous) {
        for (int l = 0; l < n; l++) {
            const int r = p264b200_stage_frame(e, step, l, fs + l);
            if (r) return r;
        }
        return P264B200_OK;
    }
    for (int l = 0; l < n; l++) {
        const int r = prepare_desc(e, step, l, fs + l);
        if (r) return r;
    }
    CK(cudaSetDevice(e->cfg.device));
    const size_t s0 = (size_t)step * e->cfg.lanes;
    CK(cudaStreamWaitEvent(e->s_h2d, e->ev_recon[step], 0));
    CK(cudaMemcpyAsync(e->d_mbs + s0 * n_mb, fs[0].mbs, (size_t)n * n_mb * sizeof(p264b200_mb), cudaMemcpyHostToDevice, e->s_h2d));
    const size_t coef_span = (size_t)(n - 1) * e->coef_cap + fs[n - 1].hdr.n_coef;
    size_t coef_used = 0;
    for (int l = 0; l < n; l++) coef_used += fs[l].hdr.n_coef;
    if (coef_span <= 2 * coef_used + (1u << 19)) {
        // tightly sized lanes (bench): one copy over the whole area
        if (coef_span)
            CK(cudaMemcpyAsync(e->d_coefs + s0 * e->coef_cap, fs[0].coefs, coef_span * sizeof(int16_t), cudaMemcpyHostToDevice, e->s_h2d));
    } else {
        // worst-case sized lanes (multi-stream decoder): the gaps would dominate the transfer, so ONE pitched copy moves the
        // used prefix of every lane (width = the longest lane's levels, one row per lane) instead of one copy per lane
        size_t longest = 0;
        for (int l = 0; l < n; l++) longest = std::max(longest, (size_t)fs[l].hdr.n_coef);
        if (longest)
            CK(cudaMemcpy2DAsync(e->d_coefs + s0 * e->coef_cap, e->coef_cap * sizeof(int16_t), fs[0].coefs, e->coef_cap * sizeof(int16_t),
                                 longest * sizeof(int16_t), n, cudaMemcpyHostToDevice, e->s_h2d));
    }
    CK(cudaMemcpyAsync(e->d_descs + s0, &e->h_descs[s0], (size_t)n * sizeof(FrameDesc), cudaMemcpyHostToDevice, e->s_h2d));
    memset(&e->desc_queued[s0], 1, n);
    CK(cudaEventRecord(e->ev_staged[step], e->s_h2d));
    e->h2d_busy = true;
    return P264B200_OK;
}

int p264b200_stage_frames_v2(p264b200_engine *e, int step, int n, const p264b200_frame_syntax_v2 *fs)
{
    if (!e || !fs || n < 1 || n > e->cfg.lanes || step < 0 || step >= e->cfg.stage_steps) return P264B200_EINVAL;
    CK(cudaSetDevice(e->cfg.device));
    const size_t n_mb = (size_t)e->g.mb_w * e->g.mb_h;
    const size_t slots = (size_t)e->cfg.stage_steps * e->cfg.lanes;
    if (!e->d_blob) {
        e->blob_cap = p264b200_pack_v2_bound(e->g.mb_w, e->g.mb_h, (uint32_t)e->coef_cap);
        CK(cudaMalloc(&e->d_blob, slots * e->blob_cap));
        CK(cudaMalloc(&e->d_v2, slots * sizeof(V2Desc)));
        CK(cudaMallocHost(&e->h_v2, slots * sizeof(V2Desc)));
    }
    // v1-shaped FrameDesc (plane pointers, slice parameters) for every lane; the records / levels it points at are
    // written by the expansion kernel below
    for (int l = 0; l < n; l++) {
        const p264b200_frame_syntax_v2 &f = fs[l];
        const bool sections_ok = f.off_hdr <= f.off_offs && f.off_offs <= f.off_mv && f.off_mv <= f.off_mask && f.off_mask <= f.off_level &&
                                 f.off_level <= f.blob_bytes && f.off_offs - f.off_hdr >= n_mb * 32 && f.off_mv - f.off_offs >= n_mb * 8 &&
                                 !((f.off_hdr | f.off_offs | f.off_mv | f.off_mask | f.off_level) & 15);
        if (!fs[l].blob || fs[l].blob_bytes > e->blob_cap || ((uintptr_t)fs[l].blob & 15) || !sections_ok) {
            set_err("p264b200_stage_frames_v2: packed picture missing, misaligned, larger than the engine's bound or with inconsistent section offsets", cudaSuccess);
            return P264B200_EINVAL;
        }
        p264b200_frame_syntax v1;
        v1.hdr = fs[l].hdr;
        v1.mbs = reinterpret_cast<const p264b200_mb *>(fs[l].blob);   // (only tested for non-NULL)
        v1.coefs = reinterpret_cast<const int16_t *>(fs[l].blob);
        const int r = prepare_desc(e, step, l, &v1);
        if (r) return r;
    }
    const size_t s0 = (size_t)step * e->cfg.lanes;
    uint8_t *dbase = e->d_blob + s0 * e->blob_cap;
    bool contiguous = true;
    size_t total = 0;
    for (int l = 0; l < n; l++) {
        if (l && fs[l].blob != fs[0].blob + total) contiguous = false;
        total += ((size_t)fs[l].blob_bytes + 15) & ~(size_t)15;
    }
    contiguous = contiguous && total <= (size_t)n * e->blob_cap;
    size_t at = 0;
    for (int l = 0; l < n; l++) {
        V2Desc &d = e->h_v2[s0 + l];
        d.blob = contiguous ? dbase + at : dbase + (size_t)l * e->blob_cap;
        d.off_hdr = fs[l].off_hdr, d.off_offs = fs[l].off_offs, d.off_mv = fs[l].off_mv, d.off_mask = fs[l].off_mask, d.off_level = fs[l].off_level;
        d.flags = fs[l].flags, d.n_coef = fs[l].hdr.n_coef, d.pad = 0;
        at += ((size_t)fs[l].blob_bytes + 15) & ~(size_t)15;
    }
    CK(cudaStreamWaitEvent(e->s_h2d, e->ev_recon[step], 0));
    if (contiguous) {
        CK(cudaMemcpyAsync(dbase, fs[0].blob, total, cudaMemcpyHostToDevice, e->s_h2d));
    } else {
        for (int l = 0; l < n; l++)
            CK(cudaMemcpyAsync(dbase + (size_t)l * e->blob_cap, fs[l].blob, fs[l].blob_bytes, cudaMemcpyHostToDevice, e->s_h2d));
    }
    CK(cudaMemcpyAsync(e->d_v2 + s0, &e->h_v2[s0], (size_t)n * sizeof(V2Desc), cudaMemcpyHostToDevice, e->s_h2d));
    CK(cudaMemcpyAsync(e->d_descs + s0, &e->h_descs[s0], (size_t)n * sizeof(FrameDesc), cudaMemcpyHostToDevice, e->s_h2d));
    memset(&e->desc_queued[s0], 1, n);
    e->launches++;
    expand_v2_kernel<<<dim3((unsigned)((n_mb + 127) / 128), n), 128, 0, e->s_h2d>>>(e->d_v2 + s0, e->d_descs + s0, (int)n_mb);
    CK(cudaGetLastError());
    CK(cudaEventRecord(e->ev_staged[step], e->s_h2d));
    e->h2d_busy = true;
    return P264B200_OK;
}

int p264b200_recon_step(p264b200_engine *e, int step, int n_lanes)
{
    if (!e || step < 0 || step >= e->cfg.stage_steps || n_lanes < 1 || n_lanes > e->cfg.lanes) return P264B200_EINVAL;
    CK(cudaSetDevice(e->cfg.device));
    const Geometry &g = e->g;
    const int n_mb = g.mb_w * g.mb_h;
    const FrameDesc *descs0 = e->d_descs + (size_t)step * e->cfg.lanes;
    const int G = n_lanes >= 2 * e->n_groups ? e->n_groups : 1;
    // G > 1: the group streams run ahead of each other ACROSS steps (no per-step barrier): a group waits for the staging
    // copy of this step, for the downloads of the slots it overwrites and for its own previous picture, nothing else --
    // so one group's deblock wavefront (ALU / latency bound) shares the SMs with the other group's recon_inter (L1 bound)
    uint32_t dst_mask = 0;   // ring slots this step overwrites: their last downloads must have read them
    for (int l = 0; l < n_lanes; l++) dst_mask |= 1u << e->slot_dst[(size_t)step * e->cfg.lanes + l];
    auto wait_inputs = [&](cudaStream_t st) -> cudaError_t {
        cudaError_t err = cudaStreamWaitEvent(st, e->ev_staged[step], 0);
        for (int sl = 0; sl < e->cfg.n_slots && err == cudaSuccess; sl++)
            if (dst_mask >> sl & 1) err = cudaStreamWaitEvent(st, e->ev_d2h_slot[sl], 0);
        return err;
    };
    if (G == 1) {
        CK(wait_inputs(e->stream));
        CK(join_groups(e));
        CK(cudaMemsetAsync(e->d_sync, 0, e->sync_bytes, e->stream));
    }
    const int words = border_threads(g.width, g.height);
    for (int gi = 0; gi < G; gi++) {
        // lanes [l0, l1) of this group; four lanes share a deblock warp, so groups start on multiples of four
        const int per = ((n_lanes + G - 1) / G + kDbfQuad - 1) & ~(kDbfQuad - 1);
        const int l0 = gi * per, l1 = l0 + per < n_lanes ? l0 + per : n_lanes;
        if (l0 >= l1) break;
        const int nl = l1 - l0;
        cudaStream_t st = G > 1 ? e->gstream[gi] : e->stream;   // recon_inter
        cudaStream_t sh = G > 1 ? e->hstream[gi] : e->stream;   // everything after it
        if (G > 1) {
            CK(wait_inputs(st));
            CK(cudaStreamWaitEvent(st, e->ev_hi_done[gi], 0));   // the group's previous picture (its reference) is finished
            // stagger: this group's MC starts when the previous group's MC is done, so that the groups do not move in phase
            if (gi > 0) CK(cudaStreamWaitEvent(st, e->ev_mc[gi - 1], 0));
        }
        unsigned flags = 0;
        for (int l = l0; l < l1; l++) flags |= e->slot_flags[(size_t)step * e->cfg.lanes + l];
        const bool intra = flags & 1, dbf = flags & 2, pslice = flags & 4;
        const FrameDesc *descs = descs0 + l0;
        int *tickets = e->d_sync + p264b200_engine::kSyncHdr * gi;
        // boundary strengths depend on the syntax alone: with one lane group the pre-pass runs on the side stream, beside
        // recon_inter (it starts once everything queued before this step is done -- the previous picture's deblock reads
        // the same strength buffer), and the main stream waits for it in front of deblock
        const bool bs_aside = dbf && pslice && G == 1 && !e->no_side;
        if (bs_aside) {
            CK(cudaEventRecord(e->ev_side_fork, st));
            CK(cudaStreamWaitEvent(e->s_side, e->ev_side_fork, 0));
            {
                ProfScope p(e, K_DEBLOCK_BS, e->s_side);
                deblock_bs_kernel<<<dim3((n_mb + 127) / 128, nl), 128, 0, e->s_side>>>(descs, g);
            }
            CK(cudaEventRecord(e->ev_side_join, e->s_side));
        }
        if (pslice) {
            ProfScope p(e, K_INTER, st);
            // CTA shape / tile height: the default and the three alternatives the measurements in DESIGN.md 7b refer to
            const int th = e->inter_variant == 2 ? 16 : e->inter_variant == 3 ? 4 : 8;
            const dim3 grid((g.mb_w + kTileW - 1) / kTileW, (g.mb_h + th - 1) / th, nl);
            switch (e->inter_variant) {
            case 1: recon_inter_kernel<512, 2, 8><<<grid, 512, sizeof(InterSmem<8>), st>>>(descs, g, e->dbg); break;
            case 2: recon_inter_kernel<576, 2, 16><<<grid, 576, sizeof(InterSmem<16>), st>>>(descs, g, e->dbg); break;
            case 3: recon_inter_kernel<192, 6, 4><<<grid, 192, sizeof(InterSmem<4>), st>>>(descs, g, e->dbg); break;
            default: recon_inter_kernel<384, 3, 8><<<grid, 384, sizeof(InterSmem<8>), st>>>(descs, g, e->dbg); break;
            }
        }
        if (G > 1) {
            CK(cudaEventRecord(e->ev_mc[gi], st));
            CK(cudaStreamWaitEvent(sh, e->ev_mc[gi], 0));
            // wavefront state of this group's lanes + its tickets
            CK(cudaMemsetAsync(e->d_sync + p264b200_engine::kSyncHdr * gi, 0, p264b200_engine::kSyncHdr * sizeof(int), sh));
            CK(cudaMemsetAsync(e->d_sync + p264b200_engine::kSyncHdr * p264b200_engine::kMaxGroups + (size_t)l0 * 3 * g.mb_h, 0, (size_t)nl * 3 * g.mb_h * sizeof(int), sh));
        }
        st = sh;
        if (bs_aside) {
            CK(cudaStreamWaitEvent(st, e->ev_side_join, 0));
        } else if (dbf) {
            ProfScope p(e, K_DEBLOCK_BS, st);
            deblock_bs_kernel<<<dim3((n_mb + 127) / 128, nl), 128, 0, st>>>(descs, g);
        }
        if (intra) {
            ProfScope p(e, K_INTRA, st);
            int max_intra = 0;
            for (int l = l0; l < l1; l++) max_intra = std::max(max_intra, e->slot_nintra[(size_t)step * e->cfg.lanes + l]);
            intra_runs_kernel<<<nl, kRunThreads, 0, st>>>(descs, g);
            recon_intra_kernel<<<(unsigned)std::min<long long>((long long)max_intra, n_mb) * nl, 32, 0, st>>>(descs, g, nl, tickets + 0, ++e->intra_epoch);
        }
        if (dbf) {
            ProfScope p(e, K_DEBLOCK, st);
            const int quads = (nl + kDbfQuad - 1) / kDbfQuad;
            const size_t dyn = G > 1 ? e->dbf_pad_bytes : 0;
            switch (e->dbf_variant) {
            case 1: deblock_kernel<6, 3><<<2 * quads * ((g.mb_h + 5) / 6), 32 * 8, dyn, st>>>(descs, g, nl, tickets, e->trace_ticket); break;
            case 2: deblock_kernel<7, 3><<<2 * quads * ((g.mb_h + 6) / 7), 32 * 9, dyn, st>>>(descs, g, nl, tickets, e->trace_ticket); break;
            default: deblock_kernel<kDbfRows, 2><<<2 * quads * ((g.mb_h + kDbfRows - 1) / kDbfRows), 32 * kDbfWarps, dyn, st>>>(descs, g, nl, tickets, e->trace_ticket); break;
            }
        }
        {
            ProfScope p(e, K_BORDER, st);
            dim3 grid((words + 255) / 256, nl);
            border_kernel<<<grid, 256, 0, st>>>(descs, g, nullptr, nullptr, nullptr);
        }
        if (G > 1) {
            CK(cudaEventRecord(e->ev_hi_done[gi], sh));
            e->groups_dirty = true;
        }
    }
    if (G > 1) CK(join_groups(e));
    CK(cudaEventRecord(e->ev_recon[step], e->stream));
    CK(cudaEventRecord(e->ev_compute, e->stream));
    CK(cudaGetLastError());
    return P264B200_OK;
}

int p264b200_recon_frame(p264b200_engine *e, int lane, const p264b200_frame_syntax *fs)
{
    // the one-picture convenience path is for the serial decoder (lane 0 only); batched callers
    // stage every lane and call p264b200_recon_step
    if (!e || lane != 0) return P264B200_EINVAL;
    int r = p264b200_stage_frame(e, 0, 0, fs);
    if (r) return r;
    return p264b200_recon_step(e, 0, 1);
}

int p264b200_frame_upload(p264b200_engine *e, int lane, int slot, const uint8_t *y, int y_stride, const uint8_t *u,
                          const uint8_t *v, int c_stride)
{
    if (!e || !y || !u || !v || lane < 0 || lane >= e->cfg.lanes || slot < 0 || slot >= e->cfg.n_slots) return P264B200_EINVAL;
    CK(cudaSetDevice(e->cfg.device));
    CK(join_groups(e));
    CK(cudaStreamWaitEvent(e->stream, e->ev_d2h_slot[slot], 0));
    const Geometry &g = e->g;
    CK(cudaMemcpy2DAsync(e->plane(lane, slot, 0), g.y_stride, y, y_stride, g.width, g.height, cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemcpy2DAsync(e->plane(lane, slot, 1), g.c_stride, u, c_stride, g.width / 2, g.height / 2, cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemcpy2DAsync(e->plane(lane, slot, 2), g.c_stride, v, c_stride, g.width / 2, g.height / 2, cudaMemcpyHostToDevice, e->stream));
    const int words = border_threads(g.width, g.height);
    e->launches++;
    border_kernel<<<dim3((words + 255) / 256, 1), 256, 0, e->stream>>>(nullptr, g, e->plane(lane, slot, 0), e->plane(lane, slot, 1),
                                                                      e->plane(lane, slot, 2));
    CK(cudaGetLastError());
    CK(cudaEventRecord(e->ev_compute, e->stream));   // downloads (s_d2h) order themselves after this upload
    return P264B200_OK;
}

int p264b200_frame_download(p264b200_engine *e, int lane, int slot, uint8_t *y, int y_stride, uint8_t *u, uint8_t *v,
                            int c_stride)
{
    if (!e || !y || !u || !v || lane < 0 || lane >= e->cfg.lanes || slot < 0 || slot >= e->cfg.n_slots) return P264B200_EINVAL;
    CK(cudaSetDevice(e->cfg.device));
    const Geometry &g = e->g;
    CK(cudaStreamWaitEvent(e->s_d2h, e->ev_compute, 0));
    CK(cudaMemcpy2DAsync(y, y_stride, e->plane(lane, slot, 0), g.y_stride, g.width, g.height, cudaMemcpyDeviceToHost, e->s_d2h));
    CK(cudaMemcpy2DAsync(u, c_stride, e->plane(lane, slot, 1), g.c_stride, g.width / 2, g.height / 2, cudaMemcpyDeviceToHost, e->s_d2h));
    CK(cudaMemcpy2DAsync(v, c_stride, e->plane(lane, slot, 2), g.c_stride, g.width / 2, g.height / 2, cudaMemcpyDeviceToHost, e->s_d2h));
    CK(cudaEventRecord(e->ev_d2h_slot[slot], e->s_d2h));
    e->d2h_busy = true;
    return P264B200_OK;
}

int p264b200_frames_download(p264b200_engine *e, int n, const int32_t *slots, uint8_t *dst, size_t picture_bytes)
{
    if (!e || !slots || !dst || n < 1 || n > e->cfg.lanes) return P264B200_EINVAL;
    const Geometry &g = e->g;
    const size_t ysz = (size_t)g.width * g.height;
    if (picture_bytes < ysz * 3 / 2 || (picture_bytes & 15) || (g.width & 31) || n > kPackMaxLanes) {
        // odd geometry: fall back to per-picture pitched copies
        if (picture_bytes < ysz * 3 / 2) return P264B200_EINVAL;
        for (int l = 0; l < n; l++) {
            uint8_t *p = dst + (size_t)l * picture_bytes;
            const int r = p264b200_frame_download(e, l, slots[l], p, g.width, p + ysz, p + ysz + ysz / 4, g.width / 2);
            if (r) return r;
        }
        return P264B200_OK;
    }
    CK(cudaSetDevice(e->cfg.device));
    const size_t need = (size_t)e->cfg.lanes * picture_bytes;
    if (need > e->out_bytes) {
        CK(cudaStreamSynchronize(e->s_d2h));
        cudaFree(e->d_out);
        e->d_out = nullptr;
        CK(cudaMalloc(&e->d_out, need));
        e->out_bytes = need;
    }
    {
        const int r = ensure_pack_table(e);
        if (r) return r;
    }
    PackSel sel;
    for (int l = 0; l < n; l++) {
        if (slots[l] < 0 || slots[l] >= e->cfg.n_slots) return P264B200_EINVAL;
        sel.slot[l] = (uint8_t)slots[l];
    }
    CK(cudaStreamWaitEvent(e->s_d2h, e->ev_compute, 0));
    e->launches++;
    pack_i420_kernel<<<dim3(64, n), 256, 0, e->s_d2h>>>(e->d_pack, e->cfg.n_slots, sel, g, e->d_out, picture_bytes);
    CK(cudaGetLastError());
    // the ring slots are free again as soon as the pack kernel has read them
    for (int l = 0; l < n; l++) CK(cudaEventRecord(e->ev_d2h_slot[slots[l]], e->s_d2h));
    CK(cudaMemcpyAsync(dst, e->d_out, (size_t)n * picture_bytes, cudaMemcpyDeviceToHost, e->s_d2h));
    e->d2h_busy = true;
    return P264B200_OK;
}

int p264b200_frame_device_planes(p264b200_engine *e, int lane, int slot, void *planes[3])
{
    if (!e || !planes || lane < 0 || lane >= e->cfg.lanes || slot < 0 || slot >= e->cfg.n_slots) return P264B200_EINVAL;
    for (int c = 0; c < 3; c++) planes[c] = e->plane(lane, slot, c);
    return P264B200_OK;
}

int p264b200_frames_md5(p264b200_engine *e, int n, const int32_t *slots, uint8_t *digests)
{
    if (!e || !slots || !digests || n < 1 || n > e->cfg.lanes || n > kPackMaxLanes) return P264B200_EINVAL;
    CK(cudaSetDevice(e->cfg.device));
    int r = ensure_pack_table(e);
    if (r) return r;
    if (!e->d_md5) CK(cudaMalloc(&e->d_md5, (size_t)e->cfg.lanes * 16));
    PackSel sel;
    for (int l = 0; l < n; l++) {
        if (slots[l] < 0 || slots[l] >= e->cfg.n_slots) return P264B200_EINVAL;
        sel.slot[l] = (uint8_t)slots[l];
    }
    CK(cudaStreamWaitEvent(e->s_d2h, e->ev_compute, 0));
    e->launches++;
    md5_i420_kernel<<<(n + 31) / 32, 32, 0, e->s_d2h>>>(e->d_pack, e->cfg.n_slots, sel, e->g, n, e->d_md5);
    CK(cudaGetLastError());
    // the ring slots are free again as soon as the kernel has read them
    for (int l = 0; l < n; l++) CK(cudaEventRecord(e->ev_d2h_slot[slots[l]], e->s_d2h));
    CK(cudaMemcpyAsync(digests, e->d_md5, (size_t)n * 16, cudaMemcpyDeviceToHost, e->s_d2h));
    e->d2h_busy = true;
    return P264B200_OK;
}

int p264b200_engine_sync(p264b200_engine *e)
{
    if (!e) return P264B200_EINVAL;
    CK(cudaSetDevice(e->cfg.device));
    CK(join_groups(e));
    CK(cudaStreamSynchronize(e->s_h2d));
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaStreamSynchronize(e->s_d2h));
    e->h2d_busy = e->d2h_busy = false;
    if (e->profile) prof_collect(e);
    return P264B200_OK;
}

void *p264b200_engine_stream(p264b200_engine *e) { return e ? (void *)e->stream : nullptr; }

int p264b200_timer_start(p264b200_engine *e)
{
    if (!e) return P264B200_EINVAL;
    CK(join_groups(e));
    CK(cudaStreamSynchronize(e->s_h2d));
    CK(cudaStreamSynchronize(e->s_d2h));
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaEventRecord(e->ev0, e->stream));
    // uploads / downloads issued from now on are ordered after the start mark
    CK(cudaStreamWaitEvent(e->s_h2d, e->ev0, 0));
    CK(cudaStreamWaitEvent(e->s_d2h, e->ev0, 0));
    return P264B200_OK;
}
int p264b200_timer_stop(p264b200_engine *e, float *ms)
{
    if (!e || !ms) return P264B200_EINVAL;
    CK(join_groups(e));
    {
        // the stop mark waits for the copy streams too, so the interval covers H2D + kernels + D2H
        cudaEvent_t a = e->ev_staged[0], b = e->ev_d2h_slot[0];
        CK(cudaEventRecord(a, e->s_h2d));
        CK(cudaEventRecord(b, e->s_d2h));
        CK(cudaStreamWaitEvent(e->stream, a, 0));
        CK(cudaStreamWaitEvent(e->stream, b, 0));
    }
    CK(cudaEventRecord(e->ev1, e->stream));
    CK(cudaEventSynchronize(e->ev1));
    CK(cudaEventElapsedTime(ms, e->ev0, e->ev1));
    if (e->profile) prof_collect(e);
    return P264B200_OK;
}

int p264b200_profile_enable(p264b200_engine *e, int on)
{
    if (!e) return P264B200_EINVAL;
    e->profile = on != 0;
    e->prof_used = 0;
    for (int i = 0; i < K_COUNT; i++) e->prof_ms[i] = 0, e->prof_n[i] = 0;
    return P264B200_OK;
}
int p264b200_profile_read(p264b200_engine *e, float ms_out[8], uint64_t launches_out[8])
{
    if (!e || !ms_out || !launches_out) return P264B200_EINVAL;
    for (int i = 0; i < 8; i++) {
        ms_out[i] = i < K_COUNT ? e->prof_ms[i] : 0.f;
        launches_out[i] = i < K_COUNT ? e->prof_n[i] : 0;
    }
    return P264B200_OK;
}
uint64_t p264b200_engine_launches(const p264b200_engine *e) { return e ? e->launches : 0; }

int p264b200_debug_cta_times(void *dst, size_t bytes)
{
    if (!dst || bytes > sizeof(g_dbf_cta_ns)) return P264B200_EINVAL;
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpyFromSymbol(dst, g_dbf_cta_ns, bytes));
    return P264B200_OK;
}

int p264b200_debug_trace(void *dst, size_t bytes)
{
    if (!dst || bytes > sizeof(g_dbf_trace)) return P264B200_EINVAL;
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpyFromSymbol(dst, g_dbf_trace, bytes));
    return P264B200_OK;
}

}  // extern "C"
