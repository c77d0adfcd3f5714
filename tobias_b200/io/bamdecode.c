/*
 * bamdecode.c -- host-side native helpers of the BAM reader (pysam / htslib are absent from the image).
 *
 *   tb_bgzf_scan      walks the BGZF block headers of a (memory-mapped) file: compressed offset / payload / uncompressed size
 *   tb_bgzf_inflate   inflates a range of blocks (zlib raw inflate) into one buffer; the Python side splits a batch over threads (ctypes
 *                     releases the GIL) -- no OpenMP here: the reader also runs in forked worker processes, where libgomp hangs
 *   tb_bam_decode     walks uncompressed alignment records and extracts the columns the hot path consumes -- the attribute set of
 *                     OneRead.from_read (tobias/utils/ngs.pyx:72-87) -- for the contigs the caller keeps:
 *       reference_start / reference_end   pos, pos + reference-consuming CIGAR ops (M, D, N, =, X)
 *       query_alignment_start / _end      leading / trailing soft clips (hard clips skipped), like pysam
 *       query_length                      l_seq, or inferred from the CIGAR (M, I, S, =, X) when l_seq == 0 (ngs.pyx:84)
 *       cigartuples[0][0], [-1][0]        first_op / last_op (255 when the record has no CIGAR)
 *     Records are validated against their block_size (a corrupt length field is an error, not an out-of-bounds read); a CIGAR of
 *     more than 65535 operations (stored in the CG:B,I tag behind a kSmN placeholder) is read from the tag.
 *
 * Build: gcc -O2 -fPIC -shared -o _bamdecode.so bamdecode.c -lz
 */
#include <stdint.h>
#include <string.h>
#include <zlib.h>

static int32_t rd32(const uint8_t *p) { int32_t v; memcpy(&v, p, 4); return v; }
static uint32_t rdu32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static uint16_t rdu16(const uint8_t *p) { uint16_t v; memcpy(&v, p, 2); return v; }

/* Block table of a BGZF file.  Returns the number of blocks (<= max_blocks), -1 on a malformed header, -2 when the table is too small. */
int64_t tb_bgzf_scan(const uint8_t *buf, int64_t n, int64_t max_blocks, int64_t *block_off, int64_t *payload_off, int32_t *payload_size,
                     int32_t *usize) {
    int64_t off = 0, k = 0;
    while (off + 18 <= n) {
        if (buf[off] != 0x1F || buf[off + 1] != 0x8B || buf[off + 2] != 8 || !(buf[off + 3] & 4)) return -1;
        const int xlen = rdu16(buf + off + 10);
        int64_t p = off + 12, bsize = -1;
        while (p + 4 <= off + 12 + xlen && p + 4 <= n) {
            const int slen = rdu16(buf + p + 2);
            if (buf[p] == 66 && buf[p + 1] == 67 && slen == 2 && p + 6 <= n) bsize = rdu16(buf + p + 4);
            p += 4 + slen;
        }
        if (bsize < 0) return -1;
        const int64_t total = bsize + 1;
        if (off + total > n) break;                                  /* truncated last block: the caller sees fewer bytes than the file holds */
        if (total < xlen + 20) return -1;
        if (k >= max_blocks) return -2;
        block_off[k] = off;
        payload_off[k] = off + 12 + xlen;
        payload_size[k] = (int32_t)(total - xlen - 20);
        usize[k] = (int32_t)rdu32(buf + off + total - 4);
        k++;
        off += total;
    }
    return k;
}

/* Inflates blocks [first, first + count) into out; out_off[i] = destination of block first + i.  Returns 0, or -(1 + index) of a bad block. */
int64_t tb_bgzf_inflate(const uint8_t *buf, const int64_t *payload_off, const int32_t *payload_size, const int32_t *usize, int64_t first,
                        int64_t count, const int64_t *out_off, uint8_t *out) {
    int64_t bad = 0;
    for (int64_t i = 0; i < count; i++) {
        const int64_t b = first + i;
        if (usize[b] == 0) continue;
        z_stream zs;
        memset(&zs, 0, sizeof(zs));
        if (inflateInit2(&zs, -15) != Z_OK) {
            bad = -(1 + b);
            continue;
        }
        zs.next_in = (Bytef *)(buf + payload_off[b]);
        zs.avail_in = (uInt)payload_size[b];
        zs.next_out = out + out_off[i];
        zs.avail_out = (uInt)usize[b];
        const int rc = inflate(&zs, Z_FINISH);
        if (rc != Z_STREAM_END || zs.avail_out != 0) bad = -(1 + b);
        inflateEnd(&zs);
    }
    return bad;
}

/* Real CIGAR of a record whose in-place CIGAR is the placeholder  <l_seq>S <ref_len>N : the CG:B,I tag.  Returns the op array or NULL. */
static const uint8_t *find_cg(const uint8_t *aux, const uint8_t *end, uint32_t *n_ops) {
    while (aux + 3 <= end) {
        const uint8_t t0 = aux[0], t1 = aux[1], ty = aux[2];
        aux += 3;
        int64_t sz;
        switch (ty) {
        case 'A': case 'c': case 'C': sz = 1; break;
        case 's': case 'S': sz = 2; break;
        case 'i': case 'I': case 'f': sz = 4; break;
        case 'Z': case 'H': {
            const uint8_t *q = aux;
            while (q < end && *q) q++;
            sz = (q - aux) + 1;
            break;
        }
        case 'B': {
            if (aux + 5 > end) return 0;
            const uint8_t sub = aux[0];
            const uint32_t cnt = rdu32(aux + 1);
            const int es = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : 4;
            if (t0 == 'C' && t1 == 'G' && sub == 'I') {
                if (aux + 5 + 4 * (int64_t)cnt > end) return 0;
                *n_ops = cnt;
                return aux + 5;
            }
            sz = 5 + (int64_t)es * cnt;
            break;
        }
        default: return 0;
        }
        if (aux + sz > end) return 0;
        aux += sz;
    }
    return 0;
}

/* Parses records from buf[0..n).  keep: per reference id 1 = extract, 0 = skip (NULL: all); records with ref_id < 0 are skipped.
 * Returns the number of records written (<= max_records), or -1 on a malformed record; *consumed receives the number of bytes of
 * complete records walked (a trailing partial record is left for the next call); *last_ref the reference id of the last record walked.
 * first_off (may be NULL; n_ref entries, -1 = not seen yet): receives base + the offset of the first record of every reference met. */
int64_t tb_bam_decode(const uint8_t *buf, int64_t n, int64_t max_records, int64_t *consumed, const uint8_t *keep, int32_t n_ref,
                      int32_t *last_ref, int64_t *first_off, int64_t base, int32_t *chrom, int32_t *start, int32_t *end, int32_t *left_soft, int32_t *right_soft,
                      int32_t *qlen, uint16_t *flag, uint8_t *first_op, uint8_t *last_op, int32_t *tlen) {
    int64_t off = 0, k = 0;
    *last_ref = -2;
    while (k < max_records && off + 4 <= n) {
        const int32_t block_size = rd32(buf + off);
        if (block_size < 32) { *consumed = off; return -1; }
        if (off + 4 + (int64_t)block_size > n) break;
        const uint8_t *r = buf + off + 4;
        const int32_t ref_id = rd32(r), pos = rd32(r + 4);
        const uint8_t l_read_name = r[8];
        uint32_t n_cigar = rdu16(r + 12);
        const uint16_t fl = rdu16(r + 14);
        const int32_t l_seq = rd32(r + 16), tl = rd32(r + 28);
        if (l_seq < 0 || 32 + (int64_t)l_read_name + 4 * (int64_t)n_cigar + ((int64_t)l_seq + 1) / 2 + l_seq > block_size) {
            *consumed = off;
            return -1;                                               /* the fixed fields do not fit the record: corrupt file */
        }
        if (first_off && ref_id >= 0 && ref_id < n_ref && first_off[ref_id] < 0) first_off[ref_id] = base + off;
        off += 4 + (int64_t)block_size;
        *last_ref = ref_id;
        if (ref_id < 0 || (keep && (ref_id >= n_ref || !keep[ref_id]))) continue;
        const uint8_t *cig = r + 32 + l_read_name;
        if (n_cigar == 2 && l_seq > 0) {                             /* kSmN placeholder: the real CIGAR sits in the CG tag */
            const uint32_t c0 = rdu32(cig), c1 = rdu32(cig + 4);
            if ((c0 & 15u) == 4 && (c0 >> 4) == (uint32_t)l_seq && (c1 & 15u) == 3) {
                uint32_t real = 0;
                const uint8_t *ops = find_cg(cig + 8 + ((int64_t)l_seq + 1) / 2 + l_seq, r + block_size, &real);
                if (!ops) { *consumed = off - 4 - block_size; return -1; }
                cig = ops;
                n_cigar = real;
            }
        }
        int64_t ref_len = 0, q_from_cigar = 0;
        int32_t ls = 0, rs = 0;
        uint8_t fo = 255, lo = 255;
        for (uint32_t i = 0; i < n_cigar; i++) {
            const uint32_t c = rdu32(cig + 4 * (int64_t)i), op = c & 15u, len = c >> 4;
            if (i == 0) fo = (uint8_t)op;
            if (i == n_cigar - 1) lo = (uint8_t)op;
            if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) ref_len += len;
            if (op == 0 || op == 1 || op == 4 || op == 7 || op == 8) q_from_cigar += len;
        }
        /* leading / trailing soft clips, skipping hard clips (pysam query_alignment_start / _end) */
        for (uint32_t i = 0; i < n_cigar; i++) {
            const uint32_t c = rdu32(cig + 4 * (int64_t)i), op = c & 15u;
            if (op == 5) continue;
            if (op == 4) ls = (int32_t)(c >> 4);
            break;
        }
        for (int64_t i = (int64_t)n_cigar - 1; i >= 0; i--) {
            const uint32_t c = rdu32(cig + 4 * i), op = c & 15u;
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
    }
    *consumed = off;
    return k;
}
