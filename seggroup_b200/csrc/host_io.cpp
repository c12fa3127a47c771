// Host-side text export of label vectors (seggroup/model.py:536-546, 564-574, 592-602: one '%d\n' line
// per raw vertex).  The consumers (kpconv/datasets/Scannet2.py:148-156, minkowski, pointgroup) read these
// files, so the format is part of the drop-in boundary; the formatting itself is plain C on the host.
#include <stdio.h>
#include <stdlib.h>
#include "../../include/seggroup_b200.h"

extern "C" int sgb_write_labels_host(const char* path, const int* values, int n) {
    if (!path || (!values && n > 0) || n < 0) return SGB_ERR_INVALID;
    FILE* f = fopen(path, "wb");
    if (!f) return SGB_ERR_INVALID;
    const size_t cap = 1 << 20;
    char* buf = (char*)malloc(cap + 16);
    if (!buf) { fclose(f); return SGB_ERR_INVALID; }
    size_t pos = 0;
    for (int i = 0; i < n; ++i) {
        long long v = values[i];
        char tmp[16];
        int len = 0;
        const bool neg = v < 0;
        if (neg) v = -v;
        do { tmp[len++] = (char)('0' + v % 10); v /= 10; } while (v);
        if (neg) buf[pos++] = '-';
        while (len) buf[pos++] = tmp[--len];
        buf[pos++] = '\n';
        if (pos >= cap) { fwrite(buf, 1, pos, f); pos = 0; }
    }
    if (pos) fwrite(buf, 1, pos, f);
    free(buf);
    return fclose(f) == 0 ? SGB_OK : SGB_ERR_INVALID;
}
