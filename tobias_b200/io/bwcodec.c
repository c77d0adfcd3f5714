/*
 * bwcodec.c -- host-side native helper of the bigWig writer (tobias_b200/io/bigwig.py; pyBigWig / libBigWig are absent from the image):
 * builds and zlib-compresses data sections / zoom blocks.  The Python side runs ranges of blocks on a thread pool (ctypes releases the GIL;
 * no OpenMP: the codec is also used in forked worker processes, where libgomp hangs).  At the benchmark's size a track holds ~10^8 non-zero
 * positions = 10^5 sections; compressing them one by one in Python took half a minute per track.
 *
 *   tb_bw_sections   variableStep sections (type 2, span 1) of <= items_per_section (position, float32 value) items
 *   tb_bw_zoom       zoom-level blocks of <= items_per_block summary records (32 bytes each)
 *
 * Both compress blocks [b0, b1) into slots of tb_bw_bound(raw block size) bytes of `out` (block b at b * slot) and report each block's
 * compressed size (-1 on a zlib error); tb_bw_pack then moves the blocks together and returns the number of bytes.
 *
 * Build: gcc -O2 -fPIC -shared -o _bwcodec.so bwcodec.c -lz
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

int64_t tb_bw_bound(int64_t raw_bytes) { return (int64_t)compressBound((uLong)raw_bytes); }

int64_t tb_bw_pack(uint8_t *out, int64_t n_blocks, int64_t slot, const int64_t *clen, int64_t *blk_off) {
    int64_t w = 0;
    for (int64_t b = 0; b < n_blocks; b++) {
        if (clen[b] < 0) return -2;
        if (w != b * slot) memmove(out + w, out + b * slot, (size_t)clen[b]);
        blk_off[b] = w;
        w += clen[b];
    }
    return w;
}

int64_t tb_bw_sections(int32_t chrom_id, const int64_t *pos, const float *val, int64_t n, int32_t items_per_section, int32_t level,
                       int64_t s0, int64_t s1, uint8_t *out, int64_t slot, int64_t *sec_len, int32_t *sec_start, int32_t *sec_end) {
    const int64_t raw_max = 24 + 8 * (int64_t)items_per_section;
    uint8_t *raw = (uint8_t *)malloc((size_t)raw_max);
    if (!raw) return -2;
    for (int64_t s = s0; s < s1; s++) {
        const int64_t a = s * items_per_section, b = a + items_per_section < n ? a + items_per_section : n;
        const uint32_t cid = (uint32_t)chrom_id, start = (uint32_t)pos[a], end = (uint32_t)pos[b - 1] + 1u, step = 0, span = 1;
        const uint16_t cnt = (uint16_t)(b - a);
        memcpy(raw, &cid, 4); memcpy(raw + 4, &start, 4); memcpy(raw + 8, &end, 4); memcpy(raw + 12, &step, 4); memcpy(raw + 16, &span, 4);
        raw[20] = 2; raw[21] = 0; memcpy(raw + 22, &cnt, 2);
        for (int64_t i = a; i < b; i++) {
            const uint32_t p = (uint32_t)pos[i];
            memcpy(raw + 24 + 8 * (i - a), &p, 4);
            memcpy(raw + 28 + 8 * (i - a), &val[i], 4);
        }
        uLongf clen = (uLongf)slot;
        const int rc = compress2(out + s * slot, &clen, raw, (uLong)(24 + 8 * (b - a)), level);
        sec_len[s] = rc == Z_OK ? (int64_t)clen : -1;
        sec_start[s] = (int32_t)start;
        sec_end[s] = (int32_t)end;
    }
    free(raw);
    return 0;
}

/* records: n x 32 bytes (chromId, start, end, validCount u32; min, max, sum, sumSquares f32), already in file order */
int64_t tb_bw_zoom(const uint8_t *records, int64_t n, int32_t items_per_block, int32_t level, int64_t b0, int64_t b1, uint8_t *out,
                   int64_t slot, int64_t *blk_len) {
    for (int64_t b = b0; b < b1; b++) {
        const int64_t a = b * items_per_block, e = a + items_per_block < n ? a + items_per_block : n;
        uLongf clen = (uLongf)slot;
        const int rc = compress2(out + b * slot, &clen, records + 32 * a, (uLong)(32 * (e - a)), level);
        blk_len[b] = rc == Z_OK ? (int64_t)clen : -1;
    }
    return 0;
}
