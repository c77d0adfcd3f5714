/*
 * hostpack.c -- host-side native helper: FASTA bases (ASCII) -> the 2-bit + N-mask layout the device genome uses
 * (tobias_b200/csrc/tb_common.cuh: base k of a contig in bits 2*(k & 15) of word k >> 4, A0 T1 C2 G3 as the reference's
 * GenomicSequence, tobias/utils/sequences.pyx:398-420; bit (k & 31) of mask word k >> 5 set for anything else).
 * A genome packed once (tobias_b200.io.fasta.PackedGenome, cached next to the FASTA like a .fai) travels to the GPU as
 * 0.375 bytes per base instead of 1.
 *
 * No OpenMP (the helpers also run in forked worker processes, where libgomp hangs): the Python side splits a contig over threads.
 *
 * Build: gcc -O3 -fPIC -shared -o _hostpack.so hostpack.c
 */
#include <stdint.h>

static inline uint32_t code_of(uint8_t c, uint32_t *is_n) {
    c &= 0xDF;                                  /* upper case */
    *is_n = 0;
    switch (c) {
    case 'A': return 0;
    case 'T': return 1;
    case 'C': return 2;
    case 'G': return 3;
    default: *is_n = 1; return 0;
    }
}

/* seq2: (n + 15) / 16 words, nmask: (n + 31) / 32 words; positions past n read as N.  Packs the groups of 32 bases [g0, g1). */
void tb_pack_ascii(const uint8_t *ascii, int64_t n, uint32_t *seq2, uint32_t *nmask, int64_t g0, int64_t g1) {
    for (int64_t g = g0; g < g1; g++) {
        uint32_t w0 = 0, w1 = 0, m = 0;
        const int64_t first = g * 32;
        for (int i = 0; i < 32; i++) {
            uint32_t isn = 1, code = 0;
            if (first + i < n) code = code_of(ascii[first + i], &isn);
            if (i < 16) w0 |= code << (2 * i); else w1 |= code << (2 * (i - 16));
            m |= isn << i;
        }
        seq2[2 * g] = w0;
        if (2 * g + 1 < (n + 15) / 16) seq2[2 * g + 1] = w1;
        nmask[g] = m;
    }
}

/* the inverse, for tests and for callers that want bases back: codes 0..3, 4 = N */
void tb_unpack_codes(const uint32_t *seq2, const uint32_t *nmask, int64_t n, uint8_t *codes) {
    for (int64_t k = 0; k < n; k++) {
        const uint32_t isn = (nmask[k >> 5] >> (k & 31)) & 1u;
        codes[k] = isn ? 4 : (uint8_t)((seq2[k >> 4] >> (2 * (k & 15))) & 3u);
    }
}
