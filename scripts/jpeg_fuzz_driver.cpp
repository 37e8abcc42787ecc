#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cstdint>
#include "stamp_b200.h"
int main(int argc, char** argv) {
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 1;
    uint32_t n;
    long ok = 0, bad = 0, total = 0;
    std::vector<int16_t> coef;
    uint16_t quant[3 * 64];
    while (fread(&n, 4, 1, f) == 1) {
        std::vector<uint8_t> buf(n);
        if (fread(buf.data(), 1, n, f) != n) break;
        // exact-size heap copy so that ASAN sees any over-read
        uint8_t* p = (uint8_t*)malloc(n ? n : 1);
        memcpy(p, buf.data(), n);
        StampJpegInfo info;
        int rc = stamp_jpeg_read_header(p, n, &info);
        if (rc == 0) {
            size_t c = stamp_jpeg_coef_count(&info);
            if (c > 0 && c < (size_t)64 * 1024 * 1024) {
                coef.assign(c, 0);
                rc = stamp_jpeg_entropy_decode(p, n, &info, coef.data(), quant);
            } else rc = -1;
        }
        (rc == 0 ? ok : bad)++;
        total++;
        free(p);
    }
    printf("inputs %ld ok %ld rejected %ld\n", total, ok, bad);
    return 0;
}
