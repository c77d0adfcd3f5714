// tb_counts.cu -- K2: cut-site weighted (di)nucleotide count tensors for the Tn5 bias model.
//
// Reference semantics: bias_estimation, tobias/tools/atacorrect_functions.py:109-181, and
// SequenceMatrix.add_sequence / add_background, tobias/utils/sequences.pyx:63-105 (PWM), :178-223 (DWM).
//   * a record belongs to every input region its alignment overlaps (fetch, :134)
//   * only records whose 5'-most CIGAR op is M count (:156-158)
//   * multiplicity n per (region, strand, cutsite), weight w = min(n, cap) (:159-168)
//   * the "within borders" test (:166) and the k-mer validity test (ngs.pyx:115) use the region extended by
//     k_flank + bg_shift and clipped to the contig (:142-143)  -- reference quirk, kept
//   * bias k-mer at the cut site, background k-mer bg_shift upstream of the read (:170-175)
//
// (1) sites: one block per (input region, tile of <= TB_K2_TILE positions) with the region's multiplicity slab IN SHARED MEMORY (u16 counters of
// both strands, 197 KB).  The records are sorted by position, so the records of a region are a contiguous range of the columns (found by a
// small search kernel).  A thread per record adds 1 to the slot of its cut site with a shared-memory atomic; the thread that finds the slot at
// zero appends the slot to a list in shared memory, so only touched slots are visited afterwards: final count -> capped weight, (region, strand,
// offset) + weight appended to the global site list, slot zeroed again (the slab is cleared once per block, not per region).  A region with
// more distinct sites than the list holds falls back to a dense scan of its slab.  (First version: a 24 GB slab in HBM for the hg38 non-peak
// space, random sector-sized read-modify-writes: 12.6 of the 16 ms of K2.)
// (2) counts from the site list: joint histograms of 2-base chunks in shared memory (default), joint histograms in L2, or one shared-memory
// atomic per position pair (PWM counts, single-chunk models) -- see below.  The reverse strand is accumulated in forward orientation and
// reverse-complemented by an index permutation on the host.  Everything is integer: bit-exact in any order.
#include "tb_common.cuh"

struct tb_reads_view2 {
    const i64 *lstart;
    const int32_t *rlen;
    const int32_t *clip;
    const uint8_t *flags;
    i64 n;
};

struct tb_k2_regions {
    const i64 *s, *e;      // original bounds (linear)
    const i64 *es, *ee;    // extended + clipped bounds (linear)
    i64 n;
};

// site record: region in the high word, strand in bit 31, offset of the cut site from the extended region start in bits 0..30
__device__ __forceinline__ u64 tb_k2_site(i64 j, int rev, i64 x) { return ((u64)j << 32) | ((u64)rev << 31) | (u64)x; }

// record range of every region: records whose alignment can overlap [s, e) start in [s - longest alignment, e)
__global__ void tb_k2_ranges_kernel(const i64 *__restrict__ lstart, i64 n_reads, tb_k2_regions R, i64 max_read_len, int sorted,
                                    i64 *__restrict__ rng) {
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 j = (i64)blockIdx.x * blockDim.x + threadIdx.x; j < R.n; j += stride) {
        i64 lo = 0, hi = n_reads;
        if (sorted) {
            const i64 v0 = R.s[j] - max_read_len, v1 = R.e[j];
            i64 a = 0, b = n_reads;
            while (a < b) {                                   // first record with lstart >= v0
                const i64 mid = (a + b) >> 1;
                if (lstart[mid] >= v0) b = mid; else a = mid + 1;
            }
            lo = a;
            b = n_reads;
            while (a < b) {                                   // first record with lstart >= v1
                const i64 mid = (a + b) >> 1;
                if (lstart[mid] >= v1) b = mid; else a = mid + 1;
            }
            hi = a;
        }
        rng[2 * j] = lo;
        rng[2 * j + 1] = hi;
    }
}

#define TB_K2_TILE 50432         // slab positions per block: a 50-kb split piece (atacorrect.py:233) + 2 x (k_flank + bg_shift) in one tile
#define TB_K2_LIST 6144          // touched slots a block can list before it falls back to the dense scan

// ctr[0] = number of sites appended, ctr[1] = overflow flag (global list capacity exceeded)
__global__ void __launch_bounds__(1024, 1)
    tb_k2_sites_kernel(tb_reads_view2 r, tb_k2_regions R, const i64 *__restrict__ rng, const int32_t *__restrict__ unit_region,
                       const int32_t *__restrict__ unit_tile, i64 n_units, int tile, int shift_fwd, int shift_rev, int cap,
                       u64 *__restrict__ sites, u32 *__restrict__ wts, i64 site_cap, unsigned long long *__restrict__ ctr,
                       u64 *__restrict__ g_scalars) {
    TB_DYN_SMEM(smem);
    const int words = tile / 2;                                // u32 words per strand (two u16 counters each)
    u32 *slab = (u32 *)smem;                                   // [2 strands][words]
    u32 *list = slab + 2 * words;                              // [TB_K2_LIST] strand << 31 | slot
    __shared__ u32 nlist;
    const unsigned FULL = 0xffffffffu;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
    for (int i = tid; i < 2 * words; i += nt) slab[i] = 0u;
    if (tid == 0) nlist = 0u;
    __syncthreads();
    long long wsum = 0;                                        // sum of the capped multiplicities this thread emitted (no_reads, :177)
    for (i64 unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        const i64 j = unit_region[unit];
        const i64 es = R.es[j], ee = R.ee[j], s0 = R.s[j], e0 = R.e[j];
        const i64 t0 = (i64)unit_tile[unit] * tile;             // first slab position of this tile, relative to es
        const i64 lo = rng[2 * j], hi = rng[2 * j + 1];
        for (i64 i = lo + tid; i < hi; i += nt) {
            const uint8_t f = r.flags[i];
            if ((f & TB_READ_SKIP) || !(f & TB_READ_5P_MATCH)) continue;
            const int rev = (f & TB_READ_REVERSE) ? 1 : 0;
            const i64 rs = r.lstart[i], re = rs + r.rlen[i];
            if (!(rs < e0 && re > s0)) continue;                // fetch: the alignment overlaps the region (:134)
            const i64 p = rev ? re + r.clip[i] + shift_rev : rs - r.clip[i] + shift_fwd;
            if (!(p >= es && p <= ee - 2)) continue;            // start < cutsite < end on the extended region (:166)
            const i64 x = p - es - t0;
            if (x < 0 || x >= tile) continue;                   // another tile of this region
            u32 *wp = slab + rev * words + (int)(x >> 1);
            const int sh = 16 * (int)(x & 1);
            const u32 old = (atomicAdd(wp, 1u << sh) >> sh) & 0xffffu;
            if (old == 0u) {
                const u32 k = atomicAdd(&nlist, 1u);
                if (k < TB_K2_LIST) list[k] = ((u32)rev << 31) | (u32)x;
            } else if (old >= 0x7fffu) {
                atomicAdd(wp, 0u - (1u << sh));                     // saturate far above any cap, far below a carry into the neighbour
            }
        }
        __syncthreads();
        const u32 n = nlist;
        if (n <= TB_K2_LIST) {
            const u32 n_round = (n + 31u) & ~31u;
            for (u32 k = tid; k < n_round; k += nt) {
                const bool have = k < n;
                u32 ent = 0, w = 0;
                if (have) {
                    ent = list[k];
                    unsigned short *cp = (unsigned short *)slab + (size_t)(ent >> 31) * (2 * words) + (ent & 0x7fffffffu);
                    const u32 cnt = *cp;
                    *cp = 0;
                    w = cnt < (u32)cap ? cnt : (u32)cap;
                    wsum += w;
                }
                const unsigned m = __ballot_sync(FULL, have);
                unsigned long long base = 0;
                const int leader = __ffs((int)m) - 1;
                if (lane == leader) base = atomicAdd(&ctr[0], (unsigned long long)__popc(m));
                base = __shfl_sync(FULL, base, leader);
                if (have) {
                    const i64 slot = (i64)base + __popc(m & ((1u << lane) - 1u));
                    if (slot < site_cap) {
                        sites[slot] = tb_k2_site(j, (int)(ent >> 31), t0 + (i64)(ent & 0x7fffffffu));
                        wts[slot] = w;
                    } else ctr[1] = 1ull;
                }
            }
        } else {                                               // more distinct sites than the list holds: dense scan of the tile
            const i64 span = ee - es - t0;
            const int used = (int)(((span < tile ? span : tile) + 1) / 2);
            for (int i = tid; i < 2 * words; i += nt) {
                const int strand = i >= words ? 1 : 0, wi = i - strand * words;
                if (wi >= used) continue;
                const u32 v = slab[i];
                if (v == 0u) continue;
                slab[i] = 0u;
                for (int h = 0; h < 2; h++) {
                    const u32 cnt = (v >> (16 * h)) & 0xffffu;
                    if (cnt == 0u) continue;
                    const u32 w = cnt < (u32)cap ? cnt : (u32)cap;
                    wsum += w;
                    const i64 slot = (i64)atomicAdd(&ctr[0], 1ull);
                    if (slot < site_cap) {
                        sites[slot] = tb_k2_site(j, strand, t0 + 2 * (i64)wi + h);
                        wts[slot] = w;
                    } else ctr[1] = 1ull;
                }
            }
        }
        __syncthreads();
        if (tid == 0) nlist = 0u;
        __syncthreads();
    }
    for (int d = 16; d > 0; d >>= 1) wsum += __shfl_xor_sync(FULL, wsum, d);
    if (lane == 0 && wsum) atomicAdd((unsigned long long *)&g_scalars[0], (unsigned long long)wsum);
}

#define TB_K2_BATCH 2048         // sites per batch of a block (2 per thread); the k-mer list holds 2 per site
#define TB_K2_PAIR_ITERS 16      // ceil(31 * 32 / 2 / 32): pair slots per lane for L <= 31

// layout of the accumulators: DWM: acc[tbl][code = Sm*4+Sn][pair, padded to a multiple of 32]   PWM: acc[tbl][S (0..4)][m]
template <int IS_DWM>
__global__ void __launch_bounds__(1024, 1)
    tb_k2_accumulate_kernel(tb_genome_view g, tb_k2_regions R, const u32 *__restrict__ wts, const u64 *__restrict__ sites, i64 n_sites,
                            int k_flank, int bg_shift, int cap, u64 *__restrict__ g_acc, u64 *__restrict__ g_scalars) {
    TB_DYN_SMEM(smem);
    const int L = 2 * k_flank + 1;
    const int npairs = L * (L + 1) / 2;
    const int pstride = (npairs + 31) & ~31;      // row stride = multiple of 32 words: lane l always hits bank l, whatever its code
    const int n_acc = IS_DWM ? 4 * 16 * pstride : 4 * 5 * L;
    int *acc = (int *)smem;
    u64 *list = (u64 *)(smem + (((size_t)n_acc * 4 + 15) & ~(size_t)15));      // [2 * TB_K2_BATCH] k-mers
    uint8_t *meta = (uint8_t *)(list + 2 * TB_K2_BATCH);                       // tbl | w << 2, 0xff = no k-mer
    __shared__ int sc[5];            // (no_reads: counted by the sites kernel), no_bias_f, no_bg_f, no_bias_r, no_bg_r
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;

    for (int i = tid; i < n_acc; i += blockDim.x) acc[i] = 0;
    if (tid < 5) sc[tid] = 0;
    // pairs owned by this lane in phase B: t = lane + 32 * it  ->  bit offsets of bases m and n in the packed k-mer
    u32 pair_sh[TB_K2_PAIR_ITERS];
    if (IS_DWM) {
#pragma unroll
        for (int it = 0; it < TB_K2_PAIR_ITERS; it++) {
            int t = lane + 32 * it, m = 0;
            pair_sh[it] = 0xffffffffu;
            if (t < npairs) {
                int rr = t;
                while (rr >= L - m) { rr -= L - m; m++; }
                pair_sh[it] = (u32)(2 * m) | ((u32)(2 * (m + rr)) << 8);
            }
        }
    }
    __syncthreads();
    i64 drained = 0;

    for (i64 b0 = (i64)blockIdx.x * TB_K2_BATCH; b0 < n_sites; b0 += (i64)gridDim.x * TB_K2_BATCH) {
        // ---- phase A: one thread per site
        for (int k = tid; k < TB_K2_BATCH; k += blockDim.x) {
            if (IS_DWM) meta[2 * k] = meta[2 * k + 1] = 0xff;
            if (b0 + k >= n_sites) continue;
            const u64 site = sites[b0 + k];
            const i64 j = (i64)(site >> 32), x = (i64)(site & 0x7fffffffu);
            const int strand = (int)((site >> 31) & 1u);
            const int w = (int)wts[b0 + k];
            const i64 es = R.es[j], ee = R.ee[j];
            const i64 gp = es + x;
            for (int which = 0; which < 2; which++) {                    // 0: bias k-mer, 1: background k-mer
                i64 c = which == 0 ? gp : (strand ? gp + bg_shift : gp - bg_shift);
                bool valid = (c > es + k_flank) && (c <= ee - k_flank - 1);          // OneRead.get_kmer, ngs.pyx:115
                u64 kmer = 0;
                u32 nf = 0xffffffffu;
                if (valid) {
                    kmer = tb_bases(g, c - k_flank, L);
                    nf = tb_nflags(g, c - k_flank, L);
                }
                int tbl = strand * 2 + which;
                if (IS_DWM) {
                    if (valid && nf == 0) {
                        list[2 * k + which] = kmer;
                        meta[2 * k + which] = (uint8_t)(tbl | (w << 2));
                        atomicAdd(&sc[1 + tbl], w);
                    }
                } else {
                    // PWM: add_sequence skips k-mers holding N (sequences.pyx:73-76); add_background does not (:100-102)
                    if (which == 0 && !(valid && nf == 0)) continue;
                    for (int m = 0; m < L; m++) {
                        int mm = strand ? L - 1 - m : m;                 // position in the read-oriented k-mer
                        int s;
                        if (!valid || ((nf >> m) & 1u)) s = 4;
                        else {
                            s = (int)((kmer >> (2 * m)) & 3u);
                            if (strand) s ^= 1;
                        }
                        atomicAdd(&acc[(tbl * 5 + s) * L + mm], w);
                    }
                    atomicAdd(&sc[1 + tbl], w);
                }
            }
        }
        if (!IS_DWM) continue;
        __syncthreads();
        // ---- phase B: one warp per k-mer, lanes over the position pairs
        for (int e = warp; e < 2 * TB_K2_BATCH; e += nwarps) {
            const int mt = meta[e];
            if (mt == 0xff) continue;
            const u64 kmer = list[e];
            const int w = mt >> 2;
            int *a = acc + (mt & 3) * 16 * pstride + lane;
#pragma unroll
            for (int it = 0; it < TB_K2_PAIR_ITERS; it++) {
                const u32 sh = pair_sh[it];
                if (sh != 0xffffffffu) {
                    int code = (int)(((kmer >> (sh & 255u)) & 3u) * 4 + ((kmer >> (sh >> 8)) & 3u));
                    atomicAdd(a + code * pstride + 32 * it, w);
                }
            }
        }
        drained += 2 * TB_K2_BATCH;
        __syncthreads();
        if (drained * cap > ((i64)1 << 30)) {                            // keep the int accumulators far from overflow
            for (int i = tid; i < n_acc; i += blockDim.x) {
                if (acc[i]) atomicAdd(&g_acc[i], (u64)(i64)acc[i]);
                acc[i] = 0;
            }
            drained = 0;
            __syncthreads();
        }
    }
    __syncthreads();
    for (int i = tid; i < n_acc; i += blockDim.x)
        if (acc[i]) atomicAdd(&g_acc[i], (u64)(i64)acc[i]);
    if (tid < 5 && sc[tid]) atomicAdd(&g_scalars[tid], (u64)(i64)sc[tid]);
}

// ------------------------------------------------------------------------------------------------ K2, joint-histogram form (DWM)
// The accumulate kernel above spends L(L+1)/2 = 325 shared-memory atomics on every k-mer.  The count tensor is a sum of
// per-position-pair dinucleotide counts, and a position pair (m, n) only needs the JOINT distribution of the two chunks of
// the k-mer that hold m and n.  So the k-mer is cut into chunks of `ch` bases (six chunks of 4 and one of 1 for L = 25), and for every pair
// of chunks the weighted joint histogram H[v_row][v_col] (4^ch x 4^ch bins, u32, L2 resident: 10 tables x 4 MB per tensor)
// receives ONE integer atomic per k-mer: 21 atomics instead of 325.  A second, tiny kernel folds every table into the
// position-pair counts (marginals over the other bases of each chunk) -- cross-chunk pairs from the table of the two
// chunks, pairs inside a chunk from the row marginal of one table in which that chunk is the row chunk (the pair {0, last}
// is stored as (last, 0) so that every chunk is a row chunk somewhere).  Integers throughout: bit-exact in any order.
#define TB_K2H_MAXT 128          // chunk pairs: L <= 31 -> 16 chunks of 2 (120 pairs, shared-memory form), 8 chunks of 4 (28 pairs, L2 form)

struct tb_k2h_desc {
    int L, ch, nc, nt, pstride;
    unsigned char row[TB_K2H_MAXT], col[TB_K2H_MAXT], within[TB_K2H_MAXT];
};

static inline bool tb_k2h_plan(int L, int ch, int pstride, tb_k2h_desc &D) {
    D.L = L; D.ch = ch; D.pstride = pstride;
    D.nc = (L + ch - 1) / ch;
    D.nt = 0;
    if (D.nc < 2) return false;
    for (int i = 0; i < D.nc; i++)
        for (int j = i + 1; j < D.nc; j++) {
            if (D.nt >= TB_K2H_MAXT) return false;
            const bool flip = (i == 0 && j == D.nc - 1 && D.nc > 2);
            D.row[D.nt] = (unsigned char)(flip ? j : i);
            D.col[D.nt] = (unsigned char)(flip ? i : j);
            D.within[D.nt] = 0;
            D.nt++;
        }
    if (D.nc == 2) {             // chunk 1 is never a row chunk: one more table, used for its row marginal only (within = 2)
        D.row[D.nt] = 1; D.col[D.nt] = 0; D.within[D.nt] = 2;
        D.nt++;
    }
    for (int c = 0; c < D.nc; c++)
        for (int t = 0; t < D.nt; t++)
            if (D.row[t] == c) {
                D.within[t] |= 1;
                break;
            }
    return true;
}

__global__ void __launch_bounds__(256)
    tb_k2_hist_kernel(tb_genome_view g, tb_k2_regions R, const u32 *__restrict__ wts, const u64 *__restrict__ sites, i64 n_sites, int k_flank,
                      int bg_shift, int cap, tb_k2h_desc D, u32 *__restrict__ H, u64 *__restrict__ g_scalars) {
    __shared__ int sc[5];            // no_reads, no_bias_f, no_bg_f, no_bias_r, no_bg_r
    if (threadIdx.x < 5) sc[threadIdx.x] = 0;
    __syncthreads();
    const int L = D.L, ch = D.ch, nt = D.nt;
    const u32 vmask = (1u << (2 * ch)) - 1u;
    const int vbits = 2 * ch;
    int mine[5] = {0, 0, 0, 0, 0};       // indexed by compile-time constants only (registers)
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x; k < n_sites; k += stride) {
        const u64 site = sites[k];
        const i64 j = (i64)(site >> 32), x = (i64)(site & 0x7fffffffu);
        const int strand = (int)((site >> 31) & 1u);
        const int w = (int)wts[k];
        const i64 es = R.es[j], ee = R.ee[j];
        const i64 gp = es + x;
#pragma unroll
        for (int which = 0; which < 2; which++) {                        // 0: bias k-mer, 1: background k-mer
            const i64 c = which == 0 ? gp : (strand ? gp + bg_shift : gp - bg_shift);
            if (!((c > es + k_flank) && (c <= ee - k_flank - 1))) continue;              // OneRead.get_kmer, ngs.pyx:115
            if (tb_nflags(g, c - k_flank, L) != 0) continue;
            const u64 kmer = tb_bases(g, c - k_flank, L);
            const int tbl = strand * 2 + which;
            mine[1 + which] += strand ? 0 : w;
            mine[3 + which] += strand ? w : 0;
            u32 *Ht = H + (((size_t)tbl * nt) << (2 * vbits));
            for (int t = 0; t < nt; t++) {
                const u32 vr = (u32)(kmer >> (vbits * D.row[t])) & vmask, vc = (u32)(kmer >> (vbits * D.col[t])) & vmask;
                atomicAdd(Ht + (((size_t)t << vbits | vr) << vbits | vc), (u32)w);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 5; i++) {
        int v = mine[i];
        for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&sc[i], v);
    }
    __syncthreads();
    if (threadIdx.x < 5 && sc[threadIdx.x]) atomicAdd(&g_scalars[threadIdx.x], (u64)(i64)sc[threadIdx.x]);
}

// ------------------------------------------------------------------------------------------------ K2, joint histograms in shared memory
// Same decomposition with chunks of TWO bases: a table has 16 x 16 bins, so the tables of every chunk pair (78 for L = 25) of both
// strands of ONE k-mer kind (bias or background) fit in the shared memory of a block: 2 x 78 x 1 KB = 156 KB.  Even blocks take the
// bias k-mer of every site, odd blocks the background k-mer; 78 shared-memory atomics per k-mer instead of 21 L2 atomics (the L2
// form above is bound by the L2 atomic rate, ~55 G/s) or 325 shared-memory atomics (accumulate kernel).  The blocks add their tables
// into the same global layout the L2 form uses and the same fold kernel finishes.
// NC > 0: number of chunks known at compile time (chunk values in registers, table offsets immediate); NC = 0: any plan
template <int NC>
__global__ void __launch_bounds__(1024, 1)
    tb_k2_hist_smem_kernel(tb_genome_view g, tb_k2_regions R, const u32 *__restrict__ wts, const u64 *__restrict__ sites, i64 n_sites,
                           int k_flank, int bg_shift, int cap, tb_k2h_desc D, u32 *__restrict__ H, u64 *__restrict__ g_scalars) {
    TB_DYN_SMEM(smem);
    u32 *T = (u32 *)smem;                                                // [strand][table][row value][column value]
    __shared__ int sc[2];                                                // weighted k-mers of this kind, forward / reverse
    const int which = blockIdx.x & 1;                                    // 0: bias k-mer, 1: background k-mer
    const int L = D.L, nt = D.nt, vbits = 2 * D.ch;
    const u32 vmask = (1u << vbits) - 1u;
    const int n_bins = (2 * nt) << (2 * vbits);
    for (int i = threadIdx.x; i < n_bins; i += blockDim.x) T[i] = 0u;
    if (threadIdx.x < 2) sc[threadIdx.x] = 0;
    __syncthreads();
    int mine_f = 0, mine_r = 0;
    const i64 stride = (i64)(gridDim.x >> 1) * blockDim.x;
    const i64 flush_every = ((i64)1 << 31) / ((i64)(cap > 0 ? cap : 1) * blockDim.x);      // per-thread k-mers before a bin could wrap
    i64 since = 0;
    for (i64 k0 = (i64)(blockIdx.x >> 1) * blockDim.x; k0 < n_sites; k0 += stride) {       // block-uniform trip count (barrier below)
        const i64 k = k0 + threadIdx.x;
        if (k < n_sites) {
            const u64 site = sites[k];
            const i64 j = (i64)(site >> 32), x = (i64)(site & 0x7fffffffu);
            const int strand = (int)((site >> 31) & 1u);
            const u32 w = wts[k];
            const i64 es = R.es[j], ee = R.ee[j];
            const i64 gp = es + x;
            const i64 c = which == 0 ? gp : (strand ? gp + bg_shift : gp - bg_shift);
            if ((c > es + k_flank) && (c <= ee - k_flank - 1) && tb_nflags(g, c - k_flank, L) == 0) {     // OneRead.get_kmer, ngs.pyx:115
                const u64 kmer = tb_bases(g, c - k_flank, L);
                mine_f += strand ? 0 : (int)w;
                mine_r += strand ? (int)w : 0;
                u32 *Tt = T + ((size_t)(strand * nt) << (2 * vbits));
                if (NC > 0) {
                    u32 v[NC > 0 ? NC : 1];
#pragma unroll
                    for (int i = 0; i < NC; i++) v[i] = (u32)(kmer >> (vbits * i)) & vmask;
                    int t = 0;
#pragma unroll
                    for (int i = 0; i < NC; i++)
#pragma unroll
                        for (int jj = i + 1; jj < NC; jj++, t++) {
                            const bool flip = (i == 0 && jj == NC - 1 && NC > 2);            // as tb_k2h_plan
                            const u32 vr = flip ? v[jj] : v[i], vc = flip ? v[i] : v[jj];
                            atomicAdd(Tt + (((size_t)t << vbits | vr) << vbits | vc), w);
                        }
                } else {
                    for (int t = 0; t < nt; t++) {
                        const u32 vr = (u32)(kmer >> (vbits * D.row[t])) & vmask, vc = (u32)(kmer >> (vbits * D.col[t])) & vmask;
                        atomicAdd(Tt + (((size_t)t << vbits | vr) << vbits | vc), w);
                    }
                }
            }
        }
        if (++since >= flush_every) {                                    // block-uniform
            __syncthreads();
            for (int i = threadIdx.x; i < n_bins; i += blockDim.x) {
                const u32 v = T[i];
                if (v) {
                    const int per = nt << (2 * vbits), st = i / per;
                    atomicAdd(&H[(size_t)(st * 2 + which) * per + (i - st * per)], v);
                    T[i] = 0u;
                }
            }
            since = 0;
            __syncthreads();
        }
    }
    for (int d = 16; d > 0; d >>= 1) {
        mine_f += __shfl_xor_sync(0xffffffffu, mine_f, d);
        mine_r += __shfl_xor_sync(0xffffffffu, mine_r, d);
    }
    if ((threadIdx.x & 31) == 0) {
        if (mine_f) atomicAdd(&sc[0], mine_f);
        if (mine_r) atomicAdd(&sc[1], mine_r);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_bins; i += blockDim.x) {
        const u32 v = T[i];
        if (v) {
            const int per = nt << (2 * vbits), st = i / per;
            atomicAdd(&H[(size_t)(st * 2 + which) * per + (i - st * per)], v);
        }
    }
    // scalars: [1] no_bias_f, [2] no_bg_f, [3] no_bias_r, [4] no_bg_r
    if (threadIdx.x < 2 && sc[threadIdx.x]) atomicAdd(&g_scalars[1 + 2 * threadIdx.x + which], (u64)(i64)sc[threadIdx.x]);
}

// Folds the joint histograms into g_acc[tbl][code = Sm*4+Sn][pair (m <= n)].  grid = (row slabs, 4 * nt); a warp per table row.
#define TB_K2H_MAXCH 5
__global__ void __launch_bounds__(256) tb_k2_fold_kernel(tb_k2h_desc D, const u32 *__restrict__ H, u64 *__restrict__ g_acc) {
    __shared__ u64 cross[TB_K2H_MAXCH * 4 * TB_K2H_MAXCH * 4];           // [(m, a)][(n, b)]
    __shared__ u64 inner[TB_K2H_MAXCH * (TB_K2H_MAXCH + 1) / 2 * 16];    // [pair inside the row chunk][code]
    __shared__ u32 wsum[8][TB_K2H_MAXCH * 4];
    const int L = D.L, ch = D.ch, vbits = 2 * ch, V = 1 << vbits;
    const int tt = blockIdx.y, tbl = tt / D.nt, t = tt % D.nt;
    const int rc = D.row[t], cc = D.col[t], flags = D.within[t];
    const int len_r = min(ch, L - rc * ch), len_c = min(ch, L - cc * ch);
    const bool do_cross = !(flags & 2), do_inner = (flags & 1) != 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int i = threadIdx.x; i < TB_K2H_MAXCH * 4 * TB_K2H_MAXCH * 4; i += blockDim.x) cross[i] = 0;
    for (int i = threadIdx.x; i < TB_K2H_MAXCH * (TB_K2H_MAXCH + 1) / 2 * 16; i += blockDim.x) inner[i] = 0;
    __syncthreads();
    const u32 *Ht = H + ((size_t)tt << (2 * vbits));
    const int rows_per_block = (V + gridDim.x - 1) / gridDim.x;
    const int r_lo = blockIdx.x * rows_per_block, r_hi = min(V, r_lo + rows_per_block);
    // lane's pair inside the row chunk (m <= n), for the inner counts
    int im = -1, in_ = -1;
    {
        int q = 0;
        for (int m = 0; m < len_r; m++)
            for (int n = m; n < len_r; n++, q++)
                if (q == lane) { im = m; in_ = n; }
    }
    for (int vr = r_lo + warp; vr < r_hi; vr += nwarps) {
        u32 Rs[TB_K2H_MAXCH][4];
#pragma unroll
        for (int n = 0; n < TB_K2H_MAXCH; n++)
#pragma unroll
            for (int b = 0; b < 4; b++) Rs[n][b] = 0u;
        const u32 *row = Ht + ((size_t)vr << vbits);
        for (int vc = lane; vc < V; vc += 32) {
            const u32 h = row[vc];
            if (h == 0u) continue;
#pragma unroll
            for (int n = 0; n < TB_K2H_MAXCH; n++) {
                const int b = (vc >> (2 * n)) & 3;
#pragma unroll
                for (int bb = 0; bb < 4; bb++) Rs[n][bb] += b == bb ? h : 0u;
            }
        }
        // warp totals through shared memory: lane n * 4 + b ends up with R[n][b]
        u32 *ws = wsum[warp];
        if (lane < TB_K2H_MAXCH * 4) ws[lane] = 0u;
        __syncwarp();
#pragma unroll
        for (int n = 0; n < TB_K2H_MAXCH; n++)
#pragma unroll
            for (int b = 0; b < 4; b++)
                if (Rs[n][b]) atomicAdd(&ws[n * 4 + b], Rs[n][b]);
        __syncwarp();
        const u32 rowsum = ws[0] + ws[1] + ws[2] + ws[3];
        const u32 my = lane < TB_K2H_MAXCH * 4 ? ws[lane] : 0u;
        __syncwarp();
        if (rowsum == 0u) continue;                                      // warp-uniform
        if (do_cross && lane < 4 * len_c && my)
            for (int m = 0; m < len_r; m++) {
                const int a = (vr >> (2 * m)) & 3;
                atomicAdd(&cross[(m * 4 + a) * (TB_K2H_MAXCH * 4) + lane], (u64)my);
            }
        if (do_inner && im >= 0) {
            const int code = ((vr >> (2 * im)) & 3) * 4 + ((vr >> (2 * in_)) & 3);
            atomicAdd(&inner[lane * 16 + code], (u64)rowsum);
        }
    }
    __syncthreads();
    u64 *ga = g_acc + (size_t)tbl * 16 * D.pstride;
    if (do_cross)
        for (int i = threadIdx.x; i < len_r * 4 * TB_K2H_MAXCH * 4; i += blockDim.x) {
            const u64 v = cross[i];
            const int ma = i / (TB_K2H_MAXCH * 4), nb = i % (TB_K2H_MAXCH * 4);
            if (!v || nb >= 4 * len_c) continue;
            int M = rc * ch + ma / 4, a = ma & 3, N = cc * ch + nb / 4, b = nb & 3;
            if (M > N) { int s = M; M = N; N = s; s = a; a = b; b = s; }
            atomicAdd(&ga[(size_t)(a * 4 + b) * D.pstride + (M * L - M * (M - 1) / 2 + (N - M))], v);
        }
    if (do_inner)
        for (int i = threadIdx.x; i < len_r * (len_r + 1) / 2 * 16; i += blockDim.x) {
            const u64 v = inner[i];
            if (!v) continue;
            int q = i / 16, code = i & 15, m = 0;
            while (q >= len_r - m) { q -= len_r - m; m++; }
            const int M = rc * ch + m, N = M + q;
            atomicAdd(&ga[(size_t)code * D.pstride + (M * L - M * (M - 1) / 2 + (N - M))], v);
        }
}

extern "C" int tb_bias_counts(tb_ctx *ctx, int64_t n_regions, const int32_t *contig, const int64_t *start, const int64_t *end,
                              int k_flank, int bg_shift, int shift_fwd, int shift_rev, int cap, int is_dwm, int64_t *counts_out,
                              int64_t *scalars_out) {
    if (!ctx || !counts_out || !scalars_out) return TB_ERR_ARG;
    if (k_flank < 0 || k_flank > 15) return tb_fail(ctx, TB_ERR_ARG, "tb_bias_counts: k_flank must be in [0, 15]");
    if (cap < 1 || cap > 63) return tb_fail(ctx, TB_ERR_ARG, "tb_bias_counts: cap must be in [1, 63]");
    if (bg_shift < 0) return tb_fail(ctx, TB_ERR_ARG, "tb_bias_counts: bg_shift must be >= 0");
    const int L = 2 * k_flank + 1;
    const int npairs = L * (L + 1) / 2;
    const int pstride = (npairs + 31) & ~31;
    const size_t n_acc = is_dwm ? (size_t)4 * 16 * pstride : (size_t)4 * 5 * L;
    const size_t n_out = is_dwm ? (size_t)4 * 25 * L * L : (size_t)4 * 5 * L;
    memset(counts_out, 0, n_out * sizeof(int64_t));
    memset(scalars_out, 0, 5 * sizeof(int64_t));
    std::vector<i64> &ls = ctx->h64[0], &le = ctx->h64[1];
    const tb_tracer tr(ctx, "tb_bias_counts");
    TB_TRY(tb_linear_regions(ctx, n_regions, contig, start, end, ls, le, false));
    if (n_regions == 0 || ctx->n_reads == 0) return TB_OK;

    const i64 ext = (i64)k_flank + bg_shift;
    std::vector<i64> &es = ctx->h64[2], &ee = ctx->h64[3];
    es.resize(n_regions);
    ee.resize(n_regions);
    for (i64 i = 0; i < n_regions; i++) {
        int c = contig[i];
        i64 lo = ctx->contig_off[c], hi = lo + ctx->contig_len[c];
        es[i] = ls[i] - ext < lo ? lo : ls[i] - ext;
        ee[i] = le[i] + ext > hi ? hi : le[i] + ext;
    }
    TB_TRY(tb_reserve(ctx, ctx->scratch1, (n_acc + 16) * sizeof(u64)));
    TB_CUDA(ctx, cudaMemsetAsync(ctx->scratch1.p, 0, (n_acc + 16) * sizeof(u64), ctx->stream));
    u64 *d_acc = ctx->scratch1.as<u64>();
    u64 *d_sc = d_acc + n_acc;
    unsigned long long *d_ctr = (unsigned long long *)(d_acc + n_acc + 8);      // [0] sites, [1] overflow, [2] read bound

    tb_reads_view2 rv{ctx->r_start.as<i64>(), ctx->r_len.as<int32_t>(), ctx->r_clip.as<int32_t>(), ctx->r_flags.as<uint8_t>(),
                      ctx->n_reads};
    tb_genome_view g{ctx->seq2.as<u32>(), ctx->nmask.as<u32>(), ctx->genome_bases};
    size_t smem = ((n_acc * 4 + 15) & ~(size_t)15) + (size_t)2 * TB_K2_BATCH * 9;
    if (is_dwm) {
        auto k = tb_k2_accumulate_kernel<1>;
        TB_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    } else {
        auto k = tb_k2_accumulate_kernel<0>;
        TB_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    i64 site_cap = ctx->n_reads + ctx->n_reads / 4 + 1024;       // one site per (record, overlapped region) at most; grown on overflow
    if (tb_opt(ctx, "k2_site_cap", 0) > 0) site_cap = tb_opt(ctx, "k2_site_cap", 0);                 // tests: force the overflow / regrow path
    if (!ctx->reads_sorted && (double)ctx->n_reads * (double)n_regions > 4.0e9)
        return tb_fail(ctx, TB_ERR_STATE, "tb_bias_counts: the records must be sorted by (contig, start) for this many regions");

    // DWM: joint-histogram form unless the model is a single chunk (options k2_no_hist / k2_chunk: comparison runs and tests)
    tb_k2h_desc HD;
    bool use_hist = false;
    if (is_dwm) {
#ifdef TB_EMU
        int ch = 3;              // the test emulator folds 64 x 64 tables (same code, fewer bins)
#else
        int ch = 4;              // 4^8 bins x 21 chunk pairs x 4 tensors = 22 MB: L2 resident (chunks of 5: 168 MB, measured slower)
#endif
        const int ev = (int)tb_opt(ctx, "k2_chunk", 0);
        if (ev >= 1 && ev <= TB_K2H_MAXCH) ch = ev;
        use_hist = tb_opt(ctx, "k2_no_hist", 0) == 0 && tb_k2h_plan(L, ch, pstride, HD);
    }
    // ... and its shared-memory form (chunks of two bases) when the tables of both strands fit in one block (L <= 29)
    tb_k2h_desc SD;
    bool use_smem_hist = false;
    size_t sh_bytes = 0;
    if (is_dwm) {
        int ch = 2;
        const int ev = (int)tb_opt(ctx, "k2_smem_chunk", 0);      // tests: 1 (more tables) or 2
        if (ev >= 1 && ev <= 2) ch = ev;
        if (tb_opt(ctx, "k2_no_hist", 0) == 0 && tb_opt(ctx, "k2_no_smem_hist", 0) == 0 && tb_k2h_plan(L, ch, pstride, SD)) {
            sh_bytes = ((size_t)2 * SD.nt << (4 * SD.ch)) * sizeof(u32);
            use_smem_hist = sh_bytes <= (size_t)200 * 1024;
        }
        if (use_smem_hist) {
            if (SD.nc == 13 && ch == 2) {
                auto k = tb_k2_hist_smem_kernel<13>;
                TB_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh_bytes));
            } else {
                auto k = tb_k2_hist_smem_kernel<0>;
                TB_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh_bytes));
            }
        }
    }

    // units of the sites kernel: (region, tile); a region longer than one tile (not produced by the ATACorrect driver: split pieces are
    // <= 50 kb) is walked tile by tile, every tile scanning the region's record range
    int tile = TB_K2_TILE;
    if (tb_opt(ctx, "k2_tile", 0) >= 64) tile = (int)(tb_opt(ctx, "k2_tile", 0) & ~63LL);             // tests: force several tiles per region
    if (tile > TB_K2_TILE) tile = TB_K2_TILE;
    std::vector<int32_t> &unit_region = ctx->h32[0], &unit_tile = ctx->h32[1];
    unit_region.clear();
    unit_tile.clear();
    unit_region.reserve(n_regions);
    unit_tile.reserve(n_regions);
    for (i64 i = 0; i < n_regions; i++) {
        const i64 len = ee[i] - es[i];
        if (len >= ((i64)1 << 31)) return tb_fail(ctx, TB_ERR_ARG, "tb_bias_counts: region %lld longer than 2^31", i);
        for (i64 t = 0; t * tile < len; t++) {
            unit_region.push_back((int32_t)i);
            unit_tile.push_back((int32_t)t);
        }
    }
    const i64 n_units = (i64)unit_region.size();
    TB_TRY(tb_h2d(ctx, ctx->reg_a, ls.data(), sizeof(i64) * n_regions));
    TB_TRY(tb_h2d(ctx, ctx->reg_b, le.data(), sizeof(i64) * n_regions));
    TB_TRY(tb_h2d(ctx, ctx->reg_c, es.data(), sizeof(i64) * n_regions));
    TB_TRY(tb_h2d(ctx, ctx->reg_d, ee.data(), sizeof(i64) * n_regions));
    TB_TRY(tb_h2d(ctx, ctx->scratch2, unit_region.data(), sizeof(int32_t) * n_units));
    TB_TRY(tb_h2d(ctx, ctx->scratch3, unit_tile.data(), sizeof(int32_t) * n_units));
    TB_TRY(tb_reserve(ctx, ctx->reg_e, sizeof(i64) * 2 * n_regions));
    tb_k2_regions R{ctx->reg_a.as<i64>(), ctx->reg_b.as<i64>(), ctx->reg_c.as<i64>(), ctx->reg_d.as<i64>(), n_regions};
    const size_t sites_smem = (size_t)tile * 2 * sizeof(unsigned short) + (size_t)TB_K2_LIST * sizeof(u32);
    {
        auto k = tb_k2_sites_kernel;
        TB_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sites_smem));
    }

    tr.mark("tables uploaded");
    tb_timer_start(ctx);
    tb_mark(ctx, 0);
    {
        i64 want = tb_div_up(n_regions, 256), capg = (i64)ctx->sm_count * 8;
        TB_LAUNCH(ctx, tb_k2_ranges_kernel, (unsigned)(want < capg ? want : capg), 256, 0, rv.lstart, rv.n, R, ctx->max_read_len,
                  ctx->reads_sorted ? 1 : 0, ctx->reg_e.as<i64>());
    }
    i64 n_sites = 0;
    for (;;) {
        TB_TRY(tb_reserve(ctx, ctx->k2_sites, (size_t)site_cap * sizeof(u64)));
        TB_TRY(tb_reserve(ctx, ctx->k2_wts, (size_t)site_cap * sizeof(u32)));
        TB_CUDA(ctx, cudaMemsetAsync(d_ctr, 0, 2 * sizeof(u64), ctx->stream));
        TB_CUDA(ctx, cudaMemsetAsync(d_sc, 0, sizeof(u64), ctx->stream));
        const unsigned grid = (unsigned)(n_units < ctx->sm_count ? n_units : ctx->sm_count);
        TB_LAUNCH(ctx, tb_k2_sites_kernel, grid, 1024, sites_smem, rv, R, (const i64 *)ctx->reg_e.as<i64>(),
                  (const int32_t *)ctx->scratch2.as<int32_t>(), (const int32_t *)ctx->scratch3.as<int32_t>(), n_units, tile, shift_fwd,
                  shift_rev, cap, ctx->k2_sites.as<u64>(), ctx->k2_wts.as<u32>(), site_cap, d_ctr, d_sc);
        unsigned long long hc[2];
        TB_TRY(tb_d2h(ctx, hc, d_ctr, sizeof(hc)));
        n_sites = (i64)hc[0];
        if (!hc[1]) break;
        site_cap = n_sites + n_sites / 8 + 1024;        // site list too small (records spanning many tiny regions): grow, redo
    }
    tb_mark(ctx, 1);
    tr.mark("sites listed (host read-back)");
    TB_TRY(tb_genome_fence(ctx, 0, ctx->n_contigs - 1));      // the k-mer kernels read the genome
    const u32 *d_wts = ctx->k2_wts.as<u32>();
    const u64 *d_sites = ctx->k2_sites.as<u64>();
    if (n_sites > 0 && use_smem_hist && (double)n_sites * cap < 4.0e9) {
        // joint histograms of 2-base chunks in shared memory (even blocks: bias k-mers, odd blocks: background k-mers), then the fold
        const size_t hbytes = ((size_t)4 * SD.nt << (4 * SD.ch)) * sizeof(u32);
        TB_TRY(tb_reserve(ctx, ctx->k2_hist, hbytes));
        TB_CUDA(ctx, cudaMemsetAsync(ctx->k2_hist.p, 0, hbytes, ctx->stream));
        i64 pairs = tb_div_up(n_sites, 1024), capp = (ctx->sm_count + 1) / 2;          // one block per SM
        if (pairs > capp) pairs = capp;
        if (SD.nc == 13 && SD.ch == 2) {
            auto k = tb_k2_hist_smem_kernel<13>;
            TB_LAUNCH(ctx, k, (unsigned)(2 * pairs), 1024, sh_bytes, g, R, d_wts, d_sites, n_sites, k_flank, bg_shift, cap, SD,
                      ctx->k2_hist.as<u32>(), d_sc);
        } else {
            auto k = tb_k2_hist_smem_kernel<0>;
            TB_LAUNCH(ctx, k, (unsigned)(2 * pairs), 1024, sh_bytes, g, R, d_wts, d_sites, n_sites, k_flank, bg_shift, cap, SD,
                      ctx->k2_hist.as<u32>(), d_sc);
        }
        const int V = 1 << (2 * SD.ch);
        const int slabs = V >= 64 ? V / 64 : 1;
        TB_LAUNCH(ctx, tb_k2_fold_kernel, dim3((unsigned)slabs, (unsigned)(4 * SD.nt)), 256, 0, SD, (const u32 *)ctx->k2_hist.as<u32>(), d_acc);
    } else if (n_sites > 0 && use_hist && (double)n_sites * cap < 4.0e9) {
        // joint-histogram form: one atomic per (k-mer, chunk pair) into L2-resident tables, then the fold
        const size_t hbytes = ((size_t)4 * HD.nt << (4 * HD.ch)) * sizeof(u32);
        TB_TRY(tb_reserve(ctx, ctx->k2_hist, hbytes));
        TB_CUDA(ctx, cudaMemsetAsync(ctx->k2_hist.p, 0, hbytes, ctx->stream));
        i64 want = tb_div_up(n_sites, 256), capg = (i64)ctx->sm_count * 8;
        TB_LAUNCH(ctx, tb_k2_hist_kernel, (unsigned)(want < capg ? want : capg), 256, 0, g, R, d_wts, d_sites, n_sites, k_flank, bg_shift, cap,
                  HD, ctx->k2_hist.as<u32>(), d_sc);
        const int V = 1 << (2 * HD.ch);
        const int slabs = V >= 64 ? V / 64 : 1;              // 64 table rows per block: 8 per warp
        TB_LAUNCH(ctx, tb_k2_fold_kernel, dim3((unsigned)slabs, (unsigned)(4 * HD.nt)), 256, 0, HD, (const u32 *)ctx->k2_hist.as<u32>(), d_acc);
    } else if (n_sites > 0) {
        i64 n_batches = tb_div_up(n_sites, TB_K2_BATCH);
        unsigned grid = (unsigned)(n_batches < ctx->sm_count ? n_batches : ctx->sm_count);
        if (is_dwm) {
            auto k = tb_k2_accumulate_kernel<1>;
            TB_LAUNCH(ctx, k, grid, 1024, smem, g, R, d_wts, d_sites, n_sites, k_flank, bg_shift, cap, d_acc, d_sc);
        } else {
            auto k = tb_k2_accumulate_kernel<0>;
            TB_LAUNCH(ctx, k, grid, 1024, smem, g, R, d_wts, d_sites, n_sites, k_flank, bg_shift, cap, d_acc, d_sc);
        }
    }
    tb_mark(ctx, 2);
    tb_timer_stop(ctx);
    tb_marks_collect(ctx, 3);      // stage_ms: [0] site list (ranges + sites kernels), [1] counts from the sites
    TB_TRY(tb_sync(ctx, "tb_bias_counts"));

    std::vector<u64> h(n_acc + 8);
    TB_TRY(tb_d2h(ctx, h.data(), d_acc, (n_acc + 8) * sizeof(u64)));
    for (int i = 0; i < 5; i++) scalars_out[i] = (int64_t)h[n_acc + i];
    if (is_dwm) {
        // expand the upper triangle (forward orientation) into [strand][bias,bg][5][5][L][L]
        std::vector<int> pair_of((size_t)L * L);
        int t = 0;
        for (int m = 0; m < L; m++)
            for (int n = m; n < L; n++) pair_of[m * L + n] = t++;
        auto T = [&](int tbl, int a, int b, int m, int n) -> int64_t {      // full symmetric forward-oriented tensor
            if (m <= n) return (int64_t)h[((size_t)tbl * 16 + a * 4 + b) * pstride + pair_of[m * L + n]];
            return (int64_t)h[((size_t)tbl * 16 + b * 4 + a) * pstride + pair_of[n * L + m]];
        };
        for (int strand = 0; strand < 2; strand++)
            for (int which = 0; which < 2; which++) {
                int tbl = strand * 2 + which;
                int64_t *dst = counts_out + (size_t)tbl * 25 * L * L;
                for (int a = 0; a < 4; a++)
                    for (int b = 0; b < 4; b++)
                        for (int m = 0; m < L; m++)
                            for (int n = 0; n < L; n++)
                                dst[(((size_t)a * 5 + b) * L + m) * L + n] =
                                    strand == 0 ? T(tbl, a, b, m, n) : T(tbl, a ^ 1, b ^ 1, L - 1 - m, L - 1 - n);
            }
    } else {
        for (size_t i = 0; i < n_acc; i++) counts_out[i] = (int64_t)h[i];
    }
    return TB_OK;
}

// ------------------------------------------------------------------------------------------------ array-level mirror of the Cython API
// SequenceMatrix.add_sequence / add_background for a batch of explicit k-mers (sequences.pyx:63-105 PWM, :178-223 DWM) -- the
// unit-level entry point the reference exposes next to the batched path above (tobias_b200.sequences uses it for its shims).
__global__ void tb_add_kmers_kernel(int L, int is_dwm, int is_background, i64 n, const i64 *__restrict__ kmers, const double *__restrict__ amounts,
                                    double *__restrict__ counts, double *__restrict__ neg, double *__restrict__ total) {
    const i64 per = is_dwm ? (i64)L * L : (i64)L;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x; t < n * per; t += stride) {
        const i64 i = t / per;
        const int r = (int)(t - i * per);
        const i64 *k = kmers + i * L;
        const double amount = amounts[i];
        bool has_n = false;
        for (int m = 0; m < L; m++) has_n = has_n || k[m] >= 4 || k[m] < 0;
        // add_sequence (both kinds) and the DWM add_background return at the first N; the PWM add_background counts it in row 4 (:100-102)
        if (has_n && (is_dwm || !is_background)) continue;
        if (is_dwm) {
            const int m = r / L, nn = r % L;
            atomicAdd(&counts[(((size_t)k[m] * 5 + k[nn]) * L + m) * L + nn], amount);
        } else {
            const int m = r;
            const int sm = (int)(k[m] < 0 || k[m] > 4 ? 4 : k[m]);
            if (is_background || amount > 0) atomicAdd(&counts[(size_t)sm * L + m], amount);
            else atomicAdd(&neg[(size_t)sm * L + m], amount);
        }
        if (r == 0) atomicAdd(total, amount);
    }
}

extern "C" int tb_add_kmers(tb_ctx *ctx, int L, int is_dwm, int is_background, int64_t n, const int64_t *kmers, const double *amounts,
                            double *counts, double *neg_counts, double *total) {
    if (!ctx || n < 0 || !counts || !total || (n > 0 && (!kmers || !amounts))) return TB_ERR_ARG;
    if (L < 1 || L > 64) return tb_fail(ctx, TB_ERR_ARG, "tb_add_kmers: L must be in [1, 64]");
    if (n == 0) return TB_OK;
    const size_t nc = is_dwm ? (size_t)25 * L * L : (size_t)5 * L;
    TB_TRY(tb_h2d(ctx, ctx->scratch2, kmers, sizeof(i64) * (size_t)n * L));
    TB_TRY(tb_h2d(ctx, ctx->scratch3, amounts, sizeof(double) * (size_t)n));
    TB_TRY(tb_reserve(ctx, ctx->scratch4, sizeof(double) * (nc + (size_t)5 * L + 1)));
    TB_CUDA(ctx, cudaMemsetAsync(ctx->scratch4.p, 0, sizeof(double) * (nc + (size_t)5 * L + 1), ctx->stream));
    double *d_counts = ctx->scratch4.as<double>(), *d_neg = d_counts + nc, *d_total = d_neg + 5 * L;
    const i64 work = (i64)n * (is_dwm ? L * L : L);
    const i64 want = tb_div_up(work, 256), capg = (i64)ctx->sm_count * 16;
    TB_LAUNCH(ctx, tb_add_kmers_kernel, (unsigned)(want < capg ? want : capg), 256, 0, L, is_dwm, is_background, (i64)n,
              (const i64 *)ctx->scratch2.as<i64>(), (const double *)ctx->scratch3.as<double>(), d_counts, d_neg, d_total);
    TB_TRY(tb_check(ctx, cudaGetLastError(), "tb_add_kmers_kernel"));
    std::vector<double> h(nc + (size_t)5 * L + 1);
    TB_TRY(tb_d2h(ctx, h.data(), d_counts, sizeof(double) * h.size()));
    for (size_t i = 0; i < nc; i++) counts[i] += h[i];
    if (neg_counts)
        for (size_t i = 0; i < (size_t)5 * L; i++) neg_counts[i] += h[nc + i];
    *total += h[nc + (size_t)5 * L];
    return TB_OK;
}
