#!/usr/bin/env python3
"""Generates small golden fixtures from the reference tree (only runs where /root/reference is
mounted; the JSON files are committed):
  tests/golden/cavlc_tables.json  -- (length, code) of every CAVLC code word, parsed from the
                                     reference's encoder-side tables core/vlc.h:32-914 (the standard's
                                     Tables 9-5, 9-7..9-10); the product's decode tables are checked
                                     against it in tests/test_host_logic.py
  tests/golden/p264_abi_layout.json -- sizeof/offsetof of the public structs of the reference's p264.h
                                     (compiled with gcc here), checked against include/p264_b200.h
"""
import json
import re
import subprocess
import sys
import tempfile
from pathlib import Path

REF = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
OUT = Path(__file__).resolve().parents[1] / "tests" / "golden"

src = (REF / "core" / "vlc.h").read_text()


def table(name):
    body = src[src.index(f"p264_{name}[") :]
    body = body[: body.index("};")]
    return [(int(a, 16), int(b)) for a, b in re.findall(r"MKVLC\(\s*(0x[0-9a-fA-F]+)\s*,\s*(\d+)\s*\)", body)]


ct = table("coeff_token")
assert len(ct) == 5 * 68
tz = table("total_zeros")
assert len(tz) == 15 * 16
tzdc = table("total_zeros_dc")
assert len(tzdc) == 12
rb = table("run_before")
assert len(rb) == 7 * 15
tables = {
    "coeff_token": [ct[i * 68 : (i + 1) * 68] for i in range(5)],  # 0..3: nC classes, 4: chroma DC
    "total_zeros": [tz[i * 16 : (i + 1) * 16] for i in range(15)],
    "total_zeros_dc": [tzdc[i * 4 : (i + 1) * 4] for i in range(3)],
    "run_before": [rb[i * 15 : (i + 1) * 15] for i in range(7)],
}
(OUT / "cavlc_tables.json").write_text(json.dumps(tables))

PROBE = r"""
#include <stdint.h>
#include <stddef.h>
#include <stdio.h>
#include HEADER
#define S(t) printf("\"sizeof " #t "\": %zu,\n", sizeof(t))
#define O(t, f) printf("\"offsetof " #t "." #f "\": %zu,\n", offsetof(t, f))
int main(void) {
  printf("{\n");
  S(p264_param_t); S(p264_image_t); S(p264_picture_t); S(p264_nal_t); S(p264_zone_t);
  O(p264_param_t, cpu); O(p264_param_t, i_csp); O(p264_param_t, vui); O(p264_param_t, i_fps_num);
  O(p264_param_t, i_frame_reference); O(p264_param_t, i_bframe); O(p264_param_t, b_deblocking_filter);
  O(p264_param_t, b_cabac); O(p264_param_t, i_cqm_preset); O(p264_param_t, psz_cqm_file); O(p264_param_t, cqm_4iy);
  O(p264_param_t, cqm_8py); O(p264_param_t, pf_log); O(p264_param_t, i_log_level); O(p264_param_t, analyse);
  O(p264_param_t, rc); O(p264_param_t, b_aud); O(p264_param_t, b_repeat_headers);
  O(p264_image_t, i_stride); O(p264_image_t, plane);
  O(p264_picture_t, i_pts); O(p264_picture_t, i_width); O(p264_picture_t, i_height); O(p264_picture_t, img);
  O(p264_nal_t, i_type); O(p264_nal_t, i_payload); O(p264_nal_t, p_payload);
  printf("\"NAL_SLICE_IDR\": %d, \"NAL_SPS\": %d, \"NAL_PPS\": %d, \"P264_CSP_I420\": %d\n}\n", NAL_SLICE_IDR, NAL_SPS, NAL_PPS, P264_CSP_I420);
  return 0;
}
"""


def probe(header):
    with tempfile.TemporaryDirectory() as td:
        c = Path(td) / "probe.c"
        c.write_text(PROBE.replace("HEADER", f'"{header}"'))
        exe = Path(td) / "probe"
        subprocess.check_call(["gcc", "-w", "-o", str(exe), str(c)])
        return json.loads(subprocess.check_output([str(exe)]).decode())


if __name__ == "__main__":
    layout = probe(str(REF / "p264.h"))
    (OUT / "p264_abi_layout.json").write_text(json.dumps(layout, indent=1))
    (OUT / "abi_probe.c.in").write_text(PROBE)
    print("ok", len(layout), "layout entries")
