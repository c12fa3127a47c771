// Host-side text export of label vectors (seggroup/model.py:536-546, 564-574, 592-602: one '%d\n' line
// per raw vertex).  The consumers (kpconv/datasets/Scannet2.py:148-156, minkowski, pointgroup) read these
// files, so the format is part of the drop-in boundary; the formatting itself is plain C on the host.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/seggroup_b200.h"

// Two digits per step from a 200-byte table; a repeated value (labels come in runs: neighbouring raw vertices mostly share a
// cluster) re-uses the previous line.  ~2 ns per line instead of ~9 for the division loop, which matters because the reference
// writes 14 files x N_raw lines per scene per forward (112 files per 8-scene batch) and the writer threads must keep up.
static const char kDigits[201] =
    "00010203040506070809101112131415161718192021222324252627282930313233343536373839404142434445464748495051525354555657585960616263646566676869"
    "707172737475767778798081828384858687888990919293949596979899";

static inline int format_line(long long v, char* out) {       // "%d\n" -> out, returns the length
    char tmp[24];
    int len = 0;
    const bool neg = v < 0;
    unsigned long long u = neg ? (unsigned long long)(-v) : (unsigned long long)v;
    while (u >= 100) { const unsigned r = (unsigned)(u % 100); u /= 100; tmp[len++] = kDigits[2 * r + 1]; tmp[len++] = kDigits[2 * r]; }
    if (u >= 10) { tmp[len++] = kDigits[2 * u + 1]; tmp[len++] = kDigits[2 * u]; }
    else tmp[len++] = (char)('0' + u);
    int pos = 0;
    if (neg) out[pos++] = '-';
    while (len) out[pos++] = tmp[--len];
    out[pos++] = '\n';
    return pos;
}

extern "C" int sgb_write_labels_host(const char* path, const int* values, int n) {
    if (!path || (!values && n > 0) || n < 0) return SGB_ERR_INVALID;
    FILE* f = fopen(path, "wb");
    if (!f) return SGB_ERR_INVALID;
    setvbuf(f, nullptr, _IONBF, 0);                               // our own 1 MB buffer below
    const size_t cap = 1 << 20;
    char* buf = (char*)malloc(cap + 32);
    if (!buf) { fclose(f); return SGB_ERR_INVALID; }
    size_t pos = 0;
    bool ok = true;
    char last[16] = {0};
    int last_len = 0, last_val = 0;
    for (int i = 0; i < n; ++i) {
        const int v = values[i];
        if (last_len == 0 || v != last_val) { last_len = format_line(v, last); last_val = v; }
        memcpy(buf + pos, last, 16);                              // one unaligned 16-byte store; only last_len bytes count
        pos += last_len;
        if (pos >= cap) { ok = ok && fwrite(buf, 1, pos, f) == pos; pos = 0; }
    }
    if (pos) ok = ok && fwrite(buf, 1, pos, f) == pos;
    free(buf);
    const bool closed = fclose(f) == 0;
    return ok && closed ? SGB_OK : SGB_ERR_INVALID;
}
