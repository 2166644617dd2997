// C-ABI of the host front-end (include/p264b200_host.h)
#include <cstdlib>
#include <cstring>
#include <new>

#include "../../../include/p264b200_host.h"
#include "cavlc.h"
#include "parser.h"

using namespace p264b200;

struct p264b200_parser {
    Parser *impl;
};

extern "C" {

p264b200_parser *p264b200_parser_open(int pinned, int verbose)
{
    p264b200_parser *p = new (std::nothrow) p264b200_parser;
    if (!p) return nullptr;
    if (pinned) {
        // page-locked buffers come from the engine side of the library; a missing device is an error
        if (p264b200_device_count() <= 0) {
            delete p;
            return nullptr;
        }
        p->impl = new (std::nothrow) Parser(p264b200_host_alloc, p264b200_host_free);
    } else
        p->impl = new (std::nothrow) Parser();
    if (!p->impl) {
        delete p;
        return nullptr;
    }
    p->impl->verbose = verbose;
    return p;
}

void p264b200_parser_close(p264b200_parser *p)
{
    if (!p) return;
    delete p->impl;
    delete p;
}

int p264b200_parser_nal(p264b200_parser *p, int nal_type, int nal_ref_idc, const uint8_t *payload, int size,
                        p264b200_frame_syntax *out, int *got_frame)
{
    if (!p || !out || !got_frame || (!payload && size > 0)) return P264B200_EINVAL;
    return p->impl->nal(nal_type, nal_ref_idc, payload, size, out, got_frame);
}

int p264b200_parser_geometry(const p264b200_parser *p, int *mb_w, int *mb_h, int *ring_size)
{
    if (!p) return P264B200_EINVAL;
    if (mb_w) *mb_w = p->impl->mb_w();
    if (mb_h) *mb_h = p->impl->mb_h();
    if (ring_size) *ring_size = p->impl->ring_size();
    return 0;
}

int p264b200_annexb_next(const uint8_t *buf, size_t size, size_t *pos, size_t *nal_start, size_t *nal_size)
{
    if (!buf || !pos || !nal_start || !nal_size) return 0;
    size_t i = *pos;
    if (i == 0) {
        // locate the first start code
        size_t z = 0;
        bool found = false;
        for (; i < size; i++) {
            if (buf[i] == 0)
                z++;
            else {
                if (buf[i] == 1 && z >= 2) {
                    found = true;
                    i++;
                    break;
                }
                z = 0;
            }
        }
        if (!found) return 0;
    }
    if (i >= size) return 0;
    const size_t start = i;
    size_t z = 0;
    for (; i < size; i++) {
        if (buf[i] == 0)
            z++;
        else {
            if (buf[i] == 1 && z >= 2) {
                *nal_start = start;
                *nal_size = i - z - start;
                *pos = i + 1;
                return 1;
            }
            z = 0;
        }
    }
    *nal_start = start;
    *nal_size = size - start;  // last NAL runs to the end of the buffer (p264decoder.c:311-317)
    *pos = size;
    return 1;
}

int p264b200_nal_unescape(const uint8_t *src, int size, uint8_t *dst, int *nal_type, int *nal_ref_idc)
{
    if (!src || !dst || !nal_type || !nal_ref_idc) return P264B200_EINVAL;
    return nal_unescape(src, size, dst, nal_type, nal_ref_idc);
}

int p264b200_cavlc_table_entry(int kind, int table, int sym, int *len, int *bits)
{
    if (!len || !bits) return P264B200_EINVAL;
    return cavlc_table_entry(kind, table, sym, len, bits);
}

}  // extern "C"
