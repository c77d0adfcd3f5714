"""Host-side logic: region algebra against the reference's RegionList (compiled, oracle/_ref), file codecs, synthetic data."""
import os
from copy import deepcopy

import numpy as np
import pytest

from tobias_b200 import synth
from tobias_b200.io import bam, bigwig, fasta
from tobias_b200.regions import OneRegion, RegionList


def _ref_regions():
    from oracle import build_ref, ref_loader
    if not (build_ref.is_built() or os.path.isdir("/root/reference/tobias")):
        pytest.skip("compiled reference not available")
    return ref_loader.load().regions


def _rand_regions(rng, n, chroms=("chr1", "chr2", "chr10")):
    out = []
    for _ in range(n):
        s = int(rng.integers(0, 5000))
        out.append([str(rng.choice(chroms)), s, s + int(rng.integers(1, 400))])
    return out


def _tuples(rl):
    return [(r.chrom, r.start, r.end) for r in rl]


def test_region_algebra_matches_reference():
    R = _ref_regions()
    rng = np.random.default_rng(0)
    for trial in range(25):
        a, b = _rand_regions(rng, 40), _rand_regions(rng, 25)
        mine_a, ref_a = RegionList([OneRegion(x) for x in a]), R.RegionList([R.OneRegion(x) for x in a])
        mine_b, ref_b = RegionList([OneRegion(x) for x in b]), R.RegionList([R.OneRegion(x) for x in b])
        mine_a.merge()
        ref_a.merge()
        assert _tuples(mine_a) == _tuples(ref_a)                      # includes the nested-interval quirk
        mine_b.merge()
        ref_b.merge()
        assert _tuples(deepcopy(mine_a).subtract(mine_b)) == _tuples(deepcopy(ref_a).subtract(ref_b))
        m2 = mine_a.apply_method(OneRegion.split_region, 97)
        r2 = ref_a.apply_method(R.OneRegion.split_region, 97)
        assert _tuples(m2) == _tuples(r2)
        info = {"chr1": 5100, "chr2": 4000}
        m3 = deepcopy(mine_a).apply_method(OneRegion.check_boundary, info, "cut")
        r3 = deepcopy(ref_a).apply_method(R.OneRegion.check_boundary, info, "cut")
        assert _tuples(m3) == _tuples(r3)
        assert [_tuples(c) for c in mine_a.chunks(7)] == [_tuples(c) for c in ref_a.chunks(7)]
        mine_a.loc_sort(["chr2", "chr10", "chr1"])
        ref_a.loc_sort(["chr2", "chr10", "chr1"])
        assert _tuples(mine_a) == _tuples(ref_a)


def test_from_bed(tmp_path):
    R = _ref_regions()
    p = tmp_path / "x.bed"
    p.write_text("#comment\nchr1\t10\t20\tname\t0\t-\nchr2\t5\t9\n")
    mine, ref = RegionList().from_bed(str(p)), R.RegionList().from_bed(str(p))
    assert _tuples(mine) == _tuples(ref) and mine[0].strand == "-" and mine[0].name == "name"
    bad = tmp_path / "bad.bed"
    bad.write_text("chr1 10 20\n")
    with pytest.raises(SystemExit):
        RegionList().from_bed(str(bad))


def test_fasta_roundtrip_with_and_without_fai(tmp_path):
    ds = synth.small_config(n_contigs=2, contig_len=30011, n_peaks=4, n_records=2000)
    path = str(tmp_path / "g.fa")
    fasta.write_fasta(path, ds.genome.items(), line_width=61)
    for use_fai in (True, False):
        if not use_fai:
            os.remove(path + ".fai")
        fa = fasta.FastaFile(path)
        assert fa.references == list(ds.genome) and fa.lengths == [len(v) for v in ds.genome.values()]
        for name, seq in ds.genome.items():
            assert np.array_equal(fa.fetch_array(name), seq)
            assert fa.fetch(name, 100, 250) == seq[100:250].tobytes().decode()
            assert np.array_equal(fa.fetch_array(name, len(seq) - 7, len(seq) + 50), seq[-7:])
        fa.close()


def test_bam_roundtrip_and_fetch_semantics(tmp_path):
    ds = synth.small_config(n_contigs=2, contig_len=40000, n_peaks=6, n_records=6000)
    path = str(tmp_path / "r.bam")
    bam.write_bam(path, ds.reads)
    rc = bam.read_bam(path)
    assert bam.read_bam_header(path) == (ds.reads.references, ds.reads.lengths)
    for f in ("chrom_id", "ref_start", "ref_end", "left_soft", "right_soft", "flag", "tlen"):
        assert np.array_equal(getattr(rc, f), getattr(ds.reads, f)), f
    ok = rc.usable()
    assert np.array_equal(rc.clip5()[ok], ds.reads.clip5()[ok]) and np.array_equal(rc.five_prime_is_match()[ok], ds.reads.five_prime_is_match()[ok])
    # fetch = htslib overlap: [ref_start, ref_end) intersects [start, end)
    idx = rc.fetch_index(0, 20000, 20100)
    brute = np.flatnonzero((rc.chrom_id == 0) & (rc.ref_start < 20100) & (rc.ref_end > 20000))
    assert np.array_equal(idx, brute)


def test_bam_reader_by_contig_indexes_and_corruption(tmp_path):
    """Bounded batches (a partial record is carried into the next batch), records of unwanted contigs skipped during the walk, and --
    with a .bai or the .tbidx sidecar -- only the blocks of the wanted contigs inflated; a corrupt length field is an error."""
    ds = synth.SynthDataset(31, {"chr1": 50000, "chr2": 42000, "chr3": 36000, "chrM": 5000}, n_peaks=12, n_records=36000)
    path = str(tmp_path / "x.bam")
    bam.write_bam(path, ds.reads, bai=True)
    full = ds.reads
    r = bam.BamReader(path)
    from_bai = r._index()
    assert from_bai is not None and from_bai[0] == r.header_end and np.all(np.diff(from_bai) > 0)
    for contigs, batch in ((["chr3"], 256 << 20), (["chr2", "chrM"], 70000), (["chr1"], 50000)):
        ids = [full.references.index(c) for c in contigs]
        m = np.isin(full.chrom_id, ids)
        got = r.read(contigs=contigs, batch_bytes=batch)
        for f in ("chrom_id", "ref_start", "ref_end", "left_soft", "right_soft", "flag", "first_op", "last_op", "tlen"):
            assert np.array_equal(getattr(got, f), getattr(full, f)[m]), (contigs, f)
    assert len(r.read(contigs=["nope"])) == 0
    r.close()
    os.remove(path + ".bai")
    r = bam.BamReader(path)
    assert not r.has_index()
    assert np.array_equal(r.read(contigs=["chr2"], batch_bytes=60000).ref_start, full.ref_start[full.chrom_id == 1])      # no index: walks from the start
    assert np.array_equal(r.build_index(batch_bytes=80000), from_bai) and os.path.exists(path + ".tbidx") and r.has_index()
    assert np.array_equal(r.read(contigs=["chr3"]).ref_end, full.ref_end[full.chrom_id == 2])
    raw = bytearray(bam.bgzf_decompress(open(path, "rb").read()))
    raw[r.header_end + 12] = raw[r.header_end + 13] = 0xFF                      # n_cigar_op of the first record: 65535 operations
    r.close()
    bad = str(tmp_path / "bad.bam")
    open(bad, "wb").write(bam.bgzf_compress(bytes(raw)))
    with pytest.raises(ValueError, match="malformed alignment record"):
        bam.read_bam(bad)
    open(bad, "wb").write(open(path, "rb").read()[:-5000])                      # truncated file
    with pytest.raises(ValueError):
        bam.read_bam(bad)


def test_bigwig_roundtrip_many_sections(tmp_path):
    rng = np.random.default_rng(1)
    path = str(tmp_path / "t.bw")
    w = bigwig.BigWigWriter(path)
    w.add_header([("chrB", 700000), ("chrA", 5000)])          # header order != sorted order
    pos = np.sort(rng.choice(700000, 400000, replace=False))
    val = rng.normal(size=pos.size).astype(np.float32)
    for a in range(0, pos.size, 50000):
        w.add_entries("chrB", pos[a:a + 50000], val[a:a + 50000])
    w.add_entries("chrA", np.arange(10, 20), np.arange(10, dtype=np.float32))
    with pytest.raises(RuntimeError):
        w.add_entries("chrA", np.arange(5, 8), np.zeros(3, np.float32))           # out of order
    w.close()
    r = bigwig.BigWigReader(path)
    assert r.chroms == {"chrB": 700000, "chrA": 5000}
    full = np.full(700000, np.nan, np.float32)
    full[pos] = val
    for s, e in ((0, 700000), (123456, 123999), (699990, 700000)):
        assert np.array_equal(r.values("chrB", s, e), full[s:e], equal_nan=True)
    assert np.array_equal(r.values("chrA", 8, 22)[2:12], np.arange(10, dtype=np.float32))
    assert r.header()["nBasesCovered"] == pos.size + 10
    with pytest.raises(RuntimeError):
        r.values("chrA", 0, 6000)
    # zoom levels (genome browsers ask for them): every level summarises every data point once; a record agrees with the data under it
    assert r.zoom_levels >= 3 and [z[0] for z in r.zoom_headers()][:3] == [64, 256, 1024]
    for lv in range(r.zoom_levels):
        red, rec = r.zoom_records(lv)
        assert rec["valid"].sum() == pos.size + 10 and np.all(rec["end"] > rec["start"]) and np.all(rec["end"] - rec["start"] <= red)
        b = rec[rec["chrom"] == 0]
        k = len(b) // 2
        seg = full[int(b["start"][k]):int(b["end"][k])]
        seg = seg[~np.isnan(seg)]
        assert b["valid"][k] == seg.size and b["min"][k] == seg.min() and b["max"][k] == seg.max() and abs(b["sum"][k] - seg.sum()) < 1e-3
    r.close()


def test_bigwig_fixedstep_sections_and_batched_reads(tmp_path):
    """Runs of consecutive positions are written as fixedStep sections (4 bytes per item), scattered ones as variableStep; both decode to
    the same values, the zoom levels do not depend on the section type, and values_many reads many intervals (abutting, overlapping,
    in gaps) exactly as values reads them one by one."""
    rng = np.random.default_rng(3)
    starts = np.sort(rng.choice(400, 120, replace=False)) * 1000
    pos = np.concatenate([np.arange(s, s + rng.integers(300, 900)) for s in starts])
    val = np.where(rng.random(pos.size) < 0.3, 0, rng.normal(size=pos.size)).astype(np.float32)
    sparse = np.sort(rng.choice(np.arange(450000, 500000), 3000, replace=False))
    sval = rng.normal(size=sparse.size).astype(np.float32)
    sizes, zooms = {}, {}
    for flag in (True, False):
        bigwig.FIXED_STEP = flag
        try:
            path = str(tmp_path / ("f%d.bw" % flag))
            w = bigwig.BigWigWriter(path)
            w.add_header([("chr1", 500000)])
            w.add_entries("chr1", pos, val)
            w.add_entries("chr1", sparse, sval)
            w.close()
        finally:
            bigwig.FIXED_STEP = True
        sizes[flag] = os.path.getsize(path)
        r = bigwig.BigWigReader(path)
        full = np.full(500000, np.nan, np.float32)
        full[pos], full[sparse] = val, sval
        assert np.array_equal(r.values("chr1", 0, 500000), full, equal_nan=True)
        zooms[flag] = r.zoom_records(0)[1]
        a = np.sort(rng.choice(499000, 300, replace=False))
        b = a + rng.integers(1, 900, size=a.size)                    # overlapping and abutting intervals among them
        flat, off = r.values_many("chr1", a, b)
        for i in range(a.size):
            assert np.array_equal(flat[off[i]:off[i + 1]], full[a[i]:b[i]], equal_nan=True), i
        flat, off = r.values_many("chr1", a[::-1], b[::-1])          # unsorted: the one-by-one path
        assert np.array_equal(flat[off[0]:off[1]], full[a[-1]:b[-1]], equal_nan=True)
        with pytest.raises(RuntimeError):
            r.values_many("chr1", [10], [600000])
        r.close()
    assert sizes[True] < 0.8 * sizes[False]
    assert np.array_equal(zooms[True], zooms[False])


def test_synthetic_data_properties():
    ds = synth.small_config()
    rc = ds.reads
    assert rc.is_sorted() and len(rc) > 50000
    assert 0.01 < ((rc.flag & 0x400) > 0).mean() < 0.04 and 0.003 < ((rc.flag & 4) > 0).mean() < 0.03
    cuts = np.where(rc.is_reverse, rc.ref_end + rc.right_soft - 5, rc.ref_start - rc.left_soft + 4)
    key = rc.chrom_id.astype(np.int64) * (1 << 32) + cuts
    assert np.unique(key[rc.usable()], return_counts=True)[1].max() > 10        # the cap-10 path is exercised
    for name, seq in ds.genome.items():
        assert (seq == ord("N")).mean() > 0.01 or len(seq) < 20000
    again = synth.small_config()
    assert np.array_equal(again.reads.ref_start, rc.ref_start) and again.peaks == ds.peaks
