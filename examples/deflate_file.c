/* deflate_file.c -- a C caller of the drop-in boundary (include/deflate_b200.h).
 *
 *   gcc -O2 -Iinclude examples/deflate_file.c -o deflate_file -Ldeflate-rs_b200 -ldeflate_b200 \
 *       -Wl,-rpath,$PWD/deflate-rs_b200
 *   ./deflate_file <in> <out> [fast|default|best] [raw|zlib|gzip]
 *
 * Does what `deflate::deflate_bytes_conf` / `_zlib_conf` / `_gzip_conf` (src/lib.rs:137,182,242) do for a
 * Rust caller: one call, host buffers in and out.  Exit status 3 = no CUDA device (the library has no CPU
 * fallback and says so instead of producing anything). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "deflate_b200.h"

int main(int argc, char **argv) {
    if (argc < 3) {
        fprintf(stderr, "usage: %s <in> <out> [fast|default|best] [raw|zlib|gzip]\n", argv[0]);
        return 2;
    }
    int preset = DFL_PRESET_DEFAULT, wrap = DFL_ZLIB;
    if (argc > 3) preset = !strcmp(argv[3], "fast") ? DFL_PRESET_FAST : !strcmp(argv[3], "best") ? DFL_PRESET_BEST : DFL_PRESET_DEFAULT;
    if (argc > 4) wrap = !strcmp(argv[4], "raw") ? DFL_RAW : !strcmp(argv[4], "gzip") ? DFL_GZIP : DFL_ZLIB;

    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    fseek(f, 0, SEEK_END);
    size_t n = (size_t)ftell(f);
    fseek(f, 0, SEEK_SET);
    uint8_t *in = (uint8_t *)malloc(n ? n : 1);
    if (fread(in, 1, n, f) != n) { fprintf(stderr, "short read\n"); return 2; }
    fclose(f);

    if (dfl_device_count() < 1) {
        fprintf(stderr, "%s\n", dfl_strerror(DFL_E_NODEVICE));
        return 3;
    }
    dfl_options opt;
    dfl_options_preset(preset, &opt);
    size_t cap = dfl_bound(n, wrap), len = 0;
    uint8_t *out = (uint8_t *)malloc(cap);
    int rc = dfl_compress(in, n, &opt, wrap, NULL, 0, out, cap, &len);
    if (rc != DFL_OK) {
        fprintf(stderr, "dfl_compress: %s (%s)\n", dfl_strerror(rc), dfl_last_cuda_error());
        return 1;
    }
    f = fopen(argv[2], "wb");
    if (!f || fwrite(out, 1, len, f) != len) { perror(argv[2]); return 2; }
    fclose(f);
    fprintf(stderr, "%zu -> %zu bytes\n", n, len);
    free(in);
    free(out);
    return 0;
}
