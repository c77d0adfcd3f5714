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

/* fixedStep sections (type 3, step = span = 1): section s holds the consecutive positions pos[sec_a[s]] .. pos[sec_b[s] - 1], 4 bytes per item */
int64_t tb_bw_sections_fixed(int32_t chrom_id, const int64_t *pos, const float *val, const int64_t *sec_a, const int64_t *sec_b, int32_t max_items,
                             int32_t level, int64_t s0, int64_t s1, uint8_t *out, int64_t slot, int64_t *sec_len, int32_t *sec_start,
                             int32_t *sec_end) {
    uint8_t *raw = (uint8_t *)malloc((size_t)(24 + 4 * (int64_t)max_items));
    if (!raw) return -2;
    for (int64_t s = s0; s < s1; s++) {
        const int64_t a = sec_a[s], b = sec_b[s];
        if (b <= a || b - a > max_items || pos[b - 1] - pos[a] != b - 1 - a) {
            sec_len[s] = -1;
            continue;
        }
        const uint32_t cid = (uint32_t)chrom_id, start = (uint32_t)pos[a], end = (uint32_t)pos[b - 1] + 1u, step = 1, span = 1;
        const uint16_t cnt = (uint16_t)(b - a);
        memcpy(raw, &cid, 4); memcpy(raw + 4, &start, 4); memcpy(raw + 8, &end, 4); memcpy(raw + 12, &step, 4); memcpy(raw + 16, &span, 4);
        raw[20] = 3; raw[21] = 0; memcpy(raw + 22, &cnt, 2);
        memcpy(raw + 24, val + a, (size_t)(4 * (b - a)));
        uLongf clen = (uLongf)slot;
        const int rc = compress2(out + s * slot, &clen, raw, (uLong)(24 + 4 * (b - a)), level);
        sec_len[s] = rc == Z_OK ? (int64_t)clen : -1;
        sec_start[s] = (int32_t)start;
        sec_end[s] = (int32_t)end;
    }
    free(raw);
    return 0;
}

/* One pass over the items of an add_entries call (positions ascending): the summaries of the first zoom level -- consecutive items whose
 * pos / red agree form one record (bin, count, min, max, sum, sum of squares) -- and the totals of the whole call in tot[0..3] = min, max,
 * sum, sum of squares.  The record arrays hold up to n entries; returns the number of records. */
int64_t tb_bw_zoom0(const int64_t *pos, const float *val, int64_t n, int32_t red, int64_t *bin, int64_t *valid, double *mn, double *mx, double *sm,
                    double *sq, double *tot) {
    int64_t r = -1, cur = -1;
    double tmin = 0.0, tmax = 0.0, tsum = 0.0, tsq = 0.0;
    for (int64_t i = 0; i < n; i++) {
        const int64_t b = pos[i] / red;
        const double v = (double)val[i];
        if (i == 0) tmin = tmax = v;
        if (v < tmin) tmin = v;
        if (v > tmax) tmax = v;
        tsum += v;
        tsq += v * v;
        if (b != cur) {
            r++;
            cur = b;
            bin[r] = b; valid[r] = 1; mn[r] = v; mx[r] = v; sm[r] = v; sq[r] = v * v;
        } else {
            valid[r]++;
            if (v < mn[r]) mn[r] = v;
            if (v > mx[r]) mx[r] = v;
            sm[r] += v;
            sq[r] += v * v;
        }
    }
    tot[0] = tmin; tot[1] = tmax; tot[2] = tsum; tot[3] = tsq;
    return r + 1;
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
