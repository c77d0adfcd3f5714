/*
 * bamdecode.c -- host-side native helper: walks uncompressed BAM alignment records and extracts the columns the hot
 * path consumes (the attribute set of OneRead.from_read, tobias/utils/ngs.pyx:72-87).  pysam / htslib are absent from
 * the image; BGZF inflation is done by the Python side (zlib), this file only parses the record stream.
 *
 *   reference_start / reference_end   pos, pos + reference-consuming CIGAR ops (M, D, N, =, X)
 *   query_alignment_start / _end      leading / trailing soft clips (hard clips skipped), like pysam
 *   query_length                      l_seq, or inferred from the CIGAR (M, I, S, =, X) when l_seq == 0 (ngs.pyx:84)
 *   cigartuples[0][0], [-1][0]        first_op / last_op (255 when the record has no CIGAR)
 *
 * Build: gcc -O2 -fPIC -shared -o _bamdecode.so bamdecode.c
 */
#include <stdint.h>
#include <string.h>

static int32_t rd32(const uint8_t *p) { int32_t v; memcpy(&v, p, 4); return v; }
static uint32_t rdu32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static uint16_t rdu16(const uint8_t *p) { uint16_t v; memcpy(&v, p, 2); return v; }

/* Parses records from buf[0..n).  Returns the number of records written (<= max_records); *consumed receives the number
 * of bytes of complete records parsed (a trailing partial record is left for the next call). */
int64_t tb_bam_decode(const uint8_t *buf, int64_t n, int64_t max_records, int64_t *consumed, int32_t *chrom, int32_t *start,
                      int32_t *end, int32_t *left_soft, int32_t *right_soft, int32_t *qlen, uint16_t *flag, uint8_t *first_op,
                      uint8_t *last_op, int32_t *tlen) {
    int64_t off = 0, k = 0;
    while (k < max_records && off + 4 <= n) {
        int32_t block_size = rd32(buf + off);
        if (block_size < 32 || off + 4 + block_size > n) break;
        const uint8_t *r = buf + off + 4;
        int32_t ref_id = rd32(r), pos = rd32(r + 4);
        uint8_t l_read_name = r[8];
        uint16_t n_cigar = rdu16(r + 12), fl = rdu16(r + 14);
        int32_t l_seq = rd32(r + 16), tl = rd32(r + 28);
        const uint8_t *cig = r + 32 + l_read_name;
        int64_t ref_len = 0, q_from_cigar = 0;
        int32_t ls = 0, rs = 0;
        uint8_t fo = 255, lo = 255;
        for (int i = 0; i < n_cigar; i++) {
            uint32_t c = rdu32(cig + 4 * (int64_t)i);
            uint32_t op = c & 15u, len = c >> 4;
            if (i == 0) fo = (uint8_t)op;
            if (i == n_cigar - 1) lo = (uint8_t)op;
            if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) ref_len += len;
            if (op == 0 || op == 1 || op == 4 || op == 7 || op == 8) q_from_cigar += len;
        }
        /* leading / trailing soft clips, skipping hard clips (pysam query_alignment_start / _end) */
        for (int i = 0; i < n_cigar; i++) {
            uint32_t c = rdu32(cig + 4 * (int64_t)i), op = c & 15u;
            if (op == 5) continue;
            if (op == 4) ls = (int32_t)(c >> 4);
            break;
        }
        for (int i = n_cigar - 1; i >= 0; i--) {
            uint32_t c = rdu32(cig + 4 * (int64_t)i), op = c & 15u;
            if (op == 5) continue;
            if (op == 4 && i > 0) rs = (int32_t)(c >> 4);
            break;
        }
        chrom[k] = ref_id;
        start[k] = pos;
        end[k] = (n_cigar == 0 || (fl & 4)) ? pos + 1 : (int32_t)(pos + ref_len);
        left_soft[k] = ls;
        right_soft[k] = rs;
        qlen[k] = l_seq > 0 ? l_seq : (int32_t)q_from_cigar;
        flag[k] = fl;
        first_op[k] = fo;
        last_op[k] = lo;
        tlen[k] = tl;
        k++;
        off += 4 + (int64_t)block_size;
    }
    *consumed = off;
    return k;
}
