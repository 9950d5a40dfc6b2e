// Error reporting, launch accounting and .flo IO for libfn2.so.
#include <stdarg.h>
#include <stdio.h>
#include <vector>
#include "fn2_common.cuh"

namespace fn2 {
static thread_local char g_err[1024] = "";
static thread_local uint64_t g_launches = 0;
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches += (uint64_t)n; }
}  // namespace fn2

extern "C" {

const char* fn2_last_error(void) { return fn2::g_err; }
const char* fn2_version(void) { return "flownet2_b200 0.1 (sm_100a)"; }
uint64_t fn2_launch_count(void) { return fn2::g_launches; }

// writeFloFile, src/caffe/util/output.cpp:44-64 ("PIEH", w, h, interleaved u,v rows)
int fn2_write_flo(const char* path, const float* flow, int h, int w) {
    FN2_CHECK_ARG(path && flow && h > 0 && w > 0, "write_flo: bad argument");
    FILE* f = fopen(path, "wb");
    if (!f) { fn2::set_error("write_flo: cannot open %s", path); return FN2_ERR_INVALID; }
    fwrite("PIEH", 1, 4, f);
    int32_t wh[2] = {w, h};
    fwrite(wh, sizeof(int32_t), 2, f);
    std::vector<float> row((size_t)w * 2);
    for (int y = 0; y < h; y++) {
        for (int x = 0; x < w; x++) {
            row[2 * x] = flow[(size_t)y * w + x];
            row[2 * x + 1] = flow[(size_t)h * w + (size_t)y * w + x];
        }
        fwrite(row.data(), sizeof(float), row.size(), f);
    }
    fclose(f);
    return FN2_OK;
}

// readFloFile, src/caffe/util/output.cpp:16-42
int fn2_read_flo(const char* path, float* flow, int* h, int* w, size_t capacity_floats) {
    FN2_CHECK_ARG(path && h && w, "read_flo: bad argument");
    FILE* f = fopen(path, "rb");
    if (!f) { fn2::set_error("read_flo: cannot open %s", path); return FN2_ERR_INVALID; }
    char tag[4];
    int32_t wh[2];
    if (fread(tag, 1, 4, f) != 4 || fread(wh, sizeof(int32_t), 2, f) != 2 || memcmp(tag, "PIEH", 4) != 0 ||
        wh[0] <= 0 || wh[1] <= 0) {
        fclose(f);
        fn2::set_error("read_flo: %s is not a .flo file", path);
        return FN2_ERR_PARSE;
    }
    *w = wh[0]; *h = wh[1];
    const size_t need = (size_t)wh[0] * wh[1] * 2;
    if (!flow) { fclose(f); return FN2_OK; }
    if (capacity_floats < need) { fclose(f); fn2::set_error("read_flo: buffer too small"); return FN2_ERR_INVALID; }
    std::vector<float> row((size_t)wh[0] * 2);
    for (int y = 0; y < wh[1]; y++) {
        if (fread(row.data(), sizeof(float), row.size(), f) != row.size()) {
            fclose(f);
            fn2::set_error("read_flo: truncated file");
            return FN2_ERR_PARSE;
        }
        for (int x = 0; x < wh[0]; x++) {
            flow[(size_t)y * wh[0] + x] = row[2 * x];
            flow[(size_t)wh[0] * wh[1] + (size_t)y * wh[0] + x] = row[2 * x + 1];
        }
    }
    fclose(f);
    return FN2_OK;
}

}  // extern "C"
