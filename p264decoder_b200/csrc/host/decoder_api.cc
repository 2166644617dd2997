// The reference's public decode API (p264.h:266,300-305,347-351,379-382) on top of the host
// syntax front-end and the GPU engine.  One handle = one stream = one engine lane; handles are
// independent (no mutable globals), like the reference.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "../../../include/p264_b200.h"
#include "../../../include/p264b200_host.h"
#include "parser.h"

using namespace p264b200;

struct p264_t {
    p264_param_t param;
    Parser *parser = nullptr;
    p264b200_engine *engine = nullptr;
    int device = 0;
    int mb_w = 0, mb_h = 0, ring = 0;
    uint8_t *mirror = nullptr;  // pinned host picture with the reference's padded geometry
    size_t mirror_bytes = 0;
    p264_picture_t pic;
};

namespace {

void default_log(void *, int level, const char *fmt, va_list ap)
{
    static const char *const names[] = {"error", "warning", "info", "debug"};
    fprintf(stderr, "p264 [%s]: ", level >= 0 && level <= 3 ? names[level] : "unknown");
    vfprintf(stderr, fmt, ap);
}

int ensure_engine(p264_t *h)
{
    int mb_w, mb_h, ring;
    mb_w = h->parser->mb_w(), mb_h = h->parser->mb_h(), ring = h->parser->ring_size();
    if (h->engine && mb_w == h->mb_w && mb_h == h->mb_h && ring == h->ring) return 0;
    if (h->engine) p264b200_engine_destroy(h->engine);
    h->engine = nullptr;
    p264b200_engine_cfg cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.device = h->device;
    cfg.lanes = 1;
    cfg.mb_w = mb_w;
    cfg.mb_h = mb_h;
    cfg.n_slots = ring;
    cfg.stage_steps = 1;
    int r = p264b200_engine_create(&h->engine, &cfg);
    if (r < 0) {
        fprintf(stderr, "p264: GPU engine creation failed (%d): %s\n", r, p264b200_last_error());
        return r;
    }
    h->mb_w = mb_w, h->mb_h = mb_h, h->ring = ring;
    // host mirror laid out like p264_frame_new (core/frame.c:42-64): stride W+64, 32/16-sample borders
    const int W = 16 * mb_w, H = 16 * mb_h, ys = W + 64, cs = ys / 2;
    const size_t need = (size_t)ys * (H + 64) + 2 * (size_t)cs * (H / 2 + 32);
    if (need > h->mirror_bytes) {
        p264b200_host_free(h->mirror);
        h->mirror = (uint8_t *)p264b200_host_alloc(need);
        if (!h->mirror) return P264B200_ENOMEM;
        h->mirror_bytes = need;
    }
    memset(&h->pic, 0, sizeof(h->pic));
    h->pic.i_width = h->param.i_width = W;
    h->pic.i_height = h->param.i_height = H;
    h->pic.img.i_csp = P264_CSP_I420;
    h->pic.img.i_plane = 3;
    h->pic.img.i_stride[0] = ys;
    h->pic.img.i_stride[1] = h->pic.img.i_stride[2] = cs;
    uint8_t *p = h->mirror;
    h->pic.img.plane[0] = p + (size_t)32 * ys + 32;
    p += (size_t)ys * (H + 64);
    h->pic.img.plane[1] = p + (size_t)16 * cs + 16;
    p += (size_t)cs * (H / 2 + 32);
    h->pic.img.plane[2] = p + (size_t)16 * cs + 16;
    return 0;
}

}  // namespace

extern "C" {

void p264_param_default(p264_param_t *param)
{
    // same observable defaults as core/core.c:41-138
    memset(param, 0, sizeof(*param));
    param->cpu = 0;  // no CPU SIMD backends here: the accelerated backend is the GPU
    param->i_threads = 1;
    param->i_csp = P264_CSP_I420;
    param->vui.i_vidformat = 5;
    param->vui.i_colorprim = 2;
    param->vui.i_transfer = 2;
    param->vui.i_colmatrix = 2;
    param->i_fps_num = 25;
    param->i_fps_den = 1;
    param->i_level_idc = 51;
    param->i_frame_reference = 1;
    param->i_keyint_max = 250;
    param->i_keyint_min = 25;
    param->i_scenecut_threshold = 40;
    param->b_bframe_adaptive = 1;
    param->b_deblocking_filter = 1;
    param->b_cabac = 1;
    param->rc.i_bitrate = 1000;
    param->rc.f_rate_tolerance = 1.0f;
    param->rc.f_vbv_buffer_init = 0.9f;
    param->rc.i_qp_constant = 26;
    param->rc.i_qp_min = 10;
    param->rc.i_qp_max = 51;
    param->rc.i_qp_step = 4;
    param->rc.f_ip_factor = 1.4f;
    param->rc.f_pb_factor = 1.3f;
    param->rc.psz_stat_out = (char *)"p264_2pass.log";
    param->rc.psz_stat_in = (char *)"p264_2pass.log";
    param->rc.psz_rc_eq = (char *)"blurCplx^(1-qComp)";
    param->rc.f_qcompress = 0.6f;
    param->rc.f_qblur = 0.5f;
    param->rc.f_complexity_blur = 20;
    param->pf_log = default_log;
    param->i_log_level = P264_LOG_INFO;
    param->analyse.intra = P264_ANALYSE_I4x4 | P264_ANALYSE_I8x8;
    param->analyse.inter = P264_ANALYSE_I4x4 | P264_ANALYSE_I8x8 | P264_ANALYSE_PSUB16x16 | P264_ANALYSE_BSUB16x16;
    param->analyse.i_direct_mv_pred = P264_DIRECT_PRED_TEMPORAL;
    param->analyse.i_me_method = P264_ME_HEX;
    param->analyse.i_me_range = 16;
    param->analyse.i_subpel_refine = 5;
    param->analyse.b_chroma_me = 1;
    param->analyse.i_mv_range = -1;
    param->analyse.b_fast_pskip = 1;
    param->analyse.b_psnr = 1;
    param->i_cqm_preset = P264_CQM_FLAT;
    memset(param->cqm_4iy, 16, 16);
    memset(param->cqm_4ic, 16, 16);
    memset(param->cqm_4py, 16, 16);
    memset(param->cqm_4pc, 16, 16);
    memset(param->cqm_8iy, 16, 64);
    memset(param->cqm_8py, 16, 64);
    param->b_repeat_headers = 1;
}

void p264_picture_alloc(p264_picture_t *pic, int i_csp, int i_width, int i_height)
{
    // core/core.c:183-251
    pic->i_type = P264_TYPE_AUTO;
    pic->i_qpplus1 = 0;
    pic->i_width = i_width;
    pic->i_height = i_height;
    pic->img.i_csp = i_csp;
    const size_t wh = (size_t)i_width * i_height;
    switch (i_csp & P264_CSP_MASK) {
    case P264_CSP_I420:
    case P264_CSP_YV12:
        pic->img.i_plane = 3;
        pic->img.plane[0] = (uint8_t *)malloc(3 * wh / 2);
        pic->img.plane[1] = pic->img.plane[0] + wh;
        pic->img.plane[2] = pic->img.plane[1] + wh / 4;
        pic->img.i_stride[0] = i_width;
        pic->img.i_stride[1] = pic->img.i_stride[2] = i_width / 2;
        break;
    case P264_CSP_I422:
        pic->img.i_plane = 3;
        pic->img.plane[0] = (uint8_t *)malloc(2 * wh);
        pic->img.plane[1] = pic->img.plane[0] + wh;
        pic->img.plane[2] = pic->img.plane[1] + wh / 2;
        pic->img.i_stride[0] = i_width;
        pic->img.i_stride[1] = pic->img.i_stride[2] = i_width / 2;
        break;
    case P264_CSP_I444:
        pic->img.i_plane = 3;
        pic->img.plane[0] = (uint8_t *)malloc(3 * wh);
        pic->img.plane[1] = pic->img.plane[0] + wh;
        pic->img.plane[2] = pic->img.plane[1] + wh;
        pic->img.i_stride[0] = pic->img.i_stride[1] = pic->img.i_stride[2] = i_width;
        break;
    case P264_CSP_YUYV:
        pic->img.i_plane = 1;
        pic->img.plane[0] = (uint8_t *)malloc(2 * wh);
        pic->img.i_stride[0] = 2 * i_width;
        break;
    case P264_CSP_RGB:
    case P264_CSP_BGR:
        pic->img.i_plane = 1;
        pic->img.plane[0] = (uint8_t *)malloc(3 * wh);
        pic->img.i_stride[0] = 3 * i_width;
        break;
    case P264_CSP_BGRA:
        pic->img.i_plane = 1;
        pic->img.plane[0] = (uint8_t *)malloc(4 * wh);
        pic->img.i_stride[0] = 4 * i_width;
        break;
    default:
        fprintf(stderr, "invalid CSP\n");
        pic->img.i_plane = 0;
        break;
    }
}

void p264_picture_clean(p264_picture_t *pic)
{
    free(pic->img.plane[0]);
    memset(pic, 0, sizeof(*pic));
}

int p264_nal_encode(void *p_data, int *pi_data, int b_annexeb, p264_nal_t *nal)
{
    // core/core.c:258-301: 4-byte start code, header byte, 00 00 03 escaping
    uint8_t *dst = (uint8_t *)p_data;
    const uint8_t *src = nal->p_payload, *end = src + nal->i_payload;
    int zeros = 0;
    if (b_annexeb) {
        *dst++ = 0, *dst++ = 0, *dst++ = 0, *dst++ = 1;
    }
    *dst++ = (uint8_t)((nal->i_ref_idc << 5) | nal->i_type);
    while (src < end) {
        if (zeros == 2 && *src <= 3) {
            *dst++ = 3;
            zeros = 0;
        }
        zeros = *src == 0 ? zeros + 1 : 0;
        *dst++ = *src++;
    }
    *pi_data = (int)(dst - (uint8_t *)p_data);
    return *pi_data;
}

int p264_nal_decode(p264_nal_t *nal, void *p_data, int i_data)
{
    int n = nal_unescape((const uint8_t *)p_data, i_data, nal->p_payload, &nal->i_type, &nal->i_ref_idc);
    if (n < 0) return -1;
    nal->i_payload = n;
    return 0;
}

p264_t *p264_decoder_open(p264_param_t *param)
{
    if (p264b200_device_count() <= 0) {
        fprintf(stderr, "p264: p264_decoder_open: no CUDA device available and this build has no CPU reconstruction path\n");
        return nullptr;
    }
    p264_t *h = new (std::nothrow) p264_t;
    if (!h) return nullptr;
    if (param)
        memcpy(&h->param, param, sizeof(*param));
    else
        p264_param_default(&h->param);
    memset(&h->pic, 0, sizeof(h->pic));
    const char *dev = getenv("P264B200_DEVICE");
    h->device = dev ? atoi(dev) : 0;
    if (h->device < 0 || h->device >= p264b200_device_count()) {
        fprintf(stderr, "p264: P264B200_DEVICE=%d out of range\n", h->device);
        delete h;
        return nullptr;
    }
    h->parser = new (std::nothrow) Parser(p264b200_host_alloc, p264b200_host_free);
    if (!h->parser) {
        delete h;
        return nullptr;
    }
    return h;
}

void p264_decoder_close(p264_t *h)
{
    if (!h) return;
    if (h->engine) p264b200_engine_destroy(h->engine);
    delete h->parser;
    p264b200_host_free(h->mirror);
    delete h;
}

int p264_decoder_decode(p264_t *h, p264_picture_t **pp_pic, p264_nal_t *nal)
{
    if (!h || !pp_pic || !nal) return -1;
    *pp_pic = nullptr;
    p264b200_frame_syntax fs;
    int got = 0;
    int r = h->parser->nal(nal->i_type, nal->i_ref_idc, nal->p_payload, nal->i_payload, &fs, &got);
    if (r < 0) return -1;
    if (!got) return 0;
    if ((r = ensure_engine(h)) < 0) return -1;
    if ((r = p264b200_recon_frame(h->engine, 0, &fs)) < 0 ||
        (r = p264b200_frame_download(h->engine, 0, fs.hdr.dst_slot, h->pic.img.plane[0], h->pic.img.i_stride[0],
                                     h->pic.img.plane[1], h->pic.img.plane[2], h->pic.img.i_stride[1])) < 0 ||
        (r = p264b200_engine_sync(h->engine)) < 0) {
        fprintf(stderr, "p264: GPU reconstruction failed (%d): %s\n", r, p264b200_last_error());
        return -1;
    }
    h->pic.i_type = fs.hdr.slice_type == P264B200_SLICE_I ? P264_TYPE_I : P264_TYPE_P;
    *pp_pic = &h->pic;
    return 0;
}

}  // extern "C"
