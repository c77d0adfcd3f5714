"""
The per-base step of `TOBIAS BINDetect` on the B200: reading footprint scores at transcription-factor binding sites and at the
random background positions of every region (reference: tobias/tools/bindetect_functions.py:266-361, `scan_and_score`).

Motif scanning itself is MOODS (third-party C++, not part of the per-base path and absent from this image), so the sites are an
input here: whatever scanner produced them, their scores are gathered from the footprint track that `score_run` left on the
device -- the track never travels to the host, only one value per site does.
"""
import random

import numpy as np

RAND_WINDOW = 200                              # bindetect_functions.py:280


def background_positions(reglen, rand_window=RAND_WINDOW):
    """Offsets inside a region at which the background distribution is sampled (bindetect_functions.py:294-297): seeded with
    the region length so that a region is processed identically regardless of its place in the file."""
    random.seed(reglen)
    return random.sample(range(reglen), max(1, int(reglen / rand_window)))


def tfbs_position(region_start, tfbs_start, tfbs_end):
    """Offset of a site's middle inside its region (bindetect_functions.py:323-324)."""
    return tfbs_start - region_start + int((tfbs_end - tfbs_start) / 2.0)


def gather_scores(ctx, region_lengths, region_starts, sites, rand_window=RAND_WINDOW):
    """Scores of the track of the last `ctx.score_run` at the sites and background positions of its regions.

    region_lengths / region_starts: trimmed length and genomic start of every scored region, in the order of the run.
    sites: iterable of (region_index, tfbs_start, tfbs_end) in genomic coordinates of the region's contig.
    Returns (site_scores float64 [n_sites], background float64 [sum of samples], background_offsets list per region);
    site scores are what scan_and_score formats with "{0:.5f}" (:331), background values what it appends to
    background_signal["signal"][condition] (:312-313), region by region.
    """
    region_lengths = np.asarray(region_lengths, dtype=np.int64)
    off = np.concatenate([[0], np.cumsum(region_lengths)])
    sites = list(sites)
    idx = np.empty(len(sites), dtype=np.int64)
    for i, (r, a, b) in enumerate(sites):
        pos = tfbs_position(int(region_starts[r]), int(a), int(b))
        if not 0 <= pos < region_lengths[r]:
            raise IndexError("site %d: middle at offset %d lies outside region %d (length %d)" % (i, pos, r, region_lengths[r]))
        idx[i] = off[r] + pos
    bg_offsets = [background_positions(int(n), rand_window) for n in region_lengths]
    bg_idx = np.concatenate([off[r] + np.asarray(p, dtype=np.int64) for r, p in enumerate(bg_offsets)]) if len(bg_offsets) else \
        np.empty(0, dtype=np.int64)
    both = ctx.score_gather(np.concatenate([idx, bg_idx]))
    return both[:len(sites)], both[len(sites):], bg_offsets
