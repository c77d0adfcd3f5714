"""
The `<prefix>_atacorrect.pdf` report of ATACorrect (reference: tobias/tools/atacorrect.py:347-348, 476-478 and the plot functions of
tobias/tools/atacorrect_functions.py:421-509): one page per strand with the Tn5 bias motif (PSSM score per position around the cut
site, one curve per nucleotide) and one page per strand with the nucleotide frequencies around the corrected cut sites before (dashed)
and after (solid) correction.

matplotlib is not a dependency of this package (it is absent from the B200 image), so the pages are drawn by a small vector-PDF writer
of our own: standard Helvetica, polylines, no compression.  The curves, colours (A green, T red, C blue, G dark khaki), axis labels,
tick labels (-flank .. flank - 1) and the cut-site line are those of the reference's figures.
"""
import numpy as np

COLORS = {0: (0.0, 0.5, 0.0), 1: (1.0, 0.0, 0.0), 2: (0.0, 0.0, 1.0), 3: (0.741, 0.718, 0.420)}      # green, red, blue, darkkhaki
NAMES = {0: "A", 1: "T", 2: "C", 3: "G"}


class _Page:
    """One page of drawing operators (PDF user space: points, origin bottom left)."""

    def __init__(self, width=640.0, height=480.0):
        self.w, self.h, self.ops = width, height, []

    def color(self, rgb, stroke=True):
        self.ops.append("%.3f %.3f %.3f %s" % (rgb[0], rgb[1], rgb[2], "RG" if stroke else "rg"))

    def line_style(self, width=1.0, dash=None):
        self.ops.append("%.2f w" % width)
        self.ops.append("[%s] 0 d" % (" ".join("%.1f" % d for d in dash) if dash else ""))

    def polyline(self, xs, ys):
        pts = [(float(x), float(y)) for x, y in zip(xs, ys) if np.isfinite(x) and np.isfinite(y)]
        if len(pts) < 2:
            return
        self.ops.append("%.2f %.2f m" % pts[0])
        self.ops += ["%.2f %.2f l" % p for p in pts[1:]]
        self.ops.append("S")

    def rect(self, x, y, w, h):
        self.ops.append("%.2f %.2f %.2f %.2f re S" % (x, y, w, h))

    def text(self, x, y, s, size=10, bold=False, align="left", rotate=False):
        s = str(s).replace("\\", "\\\\").replace("(", "\\(").replace(")", "\\)")
        width = 0.5 * size * len(s)                                  # Helvetica averages about half an em per character
        if align == "center":
            x, y = (x, y - width / 2.0) if rotate else (x - width / 2.0, y)
        elif align == "right":
            x -= width
        self.color((0, 0, 0), stroke=False)
        m = "0 1 -1 0 %.2f %.2f Tm" % (x, y) if rotate else "1 0 0 1 %.2f %.2f Tm" % (x, y)
        self.ops.append("BT /%s %d Tf %s (%s) Tj ET" % ("F2" if bold else "F1", size, m, s))

    def stream(self):
        return "\n".join(self.ops).encode("latin-1", "replace")


def write_pdf(path, pages):
    objs = []                                                        # (object number, bytes)

    def add(body):
        objs.append(body)
        return len(objs)
    font1 = add(b"<< /Type /Font /Subtype /Type1 /BaseFont /Helvetica >>")
    font2 = add(b"<< /Type /Font /Subtype /Type1 /BaseFont /Helvetica-Bold >>")
    pages_id = len(objs) + 2 * len(pages) + 1
    kids = []
    for pg in pages:
        data = pg.stream()
        content = add(b"<< /Length %d >>\nstream\n" % len(data) + data + b"\nendstream")
        kids.append(add(("<< /Type /Page /Parent %d 0 R /MediaBox [0 0 %.0f %.0f] /Contents %d 0 R /Resources << /Font << /F1 %d 0 R /F2 %d 0 R >> >> >>"
                         % (pages_id, pg.w, pg.h, content, font1, font2)).encode()))
    assert len(objs) + 1 == pages_id
    add(("<< /Type /Pages /Kids [%s] /Count %d >>" % (" ".join("%d 0 R" % k for k in kids), len(kids))).encode())
    catalog = add(("<< /Type /Catalog /Pages %d 0 R >>" % pages_id).encode())
    out = bytearray(b"%PDF-1.4\n")
    offsets = []
    for i, body in enumerate(objs):
        offsets.append(len(out))
        out += b"%d 0 obj\n" % (i + 1) + body + b"\nendobj\n"
    xref = len(out)
    out += b"xref\n0 %d\n0000000000 65535 f \n" % (len(objs) + 1)
    for off in offsets:
        out += b"%010d 00000 n \n" % off
    out += b"trailer\n<< /Size %d /Root %d 0 R >>\nstartxref\n%d\n%%%%EOF\n" % (len(objs) + 1, catalog, xref)
    with open(path, "wb") as fh:
        fh.write(bytes(out))


def _nice_ticks(lo, hi, n=6):
    if not np.isfinite(lo) or not np.isfinite(hi) or hi <= lo:
        lo, hi = (lo - 1.0, lo + 1.0) if np.isfinite(lo) else (0.0, 1.0)
    step = 10.0 ** np.floor(np.log10((hi - lo) / n))
    for mult in (1.0, 2.0, 5.0, 10.0):
        if (hi - lo) / (step * mult) <= n:
            step *= mult
            break
    first = np.ceil(lo / step) * step
    return [float(v) for v in np.arange(first, hi + step * 1e-9, step)]


def _axes(pg, title, xlabel, ylabel, length, ymin, ymax):
    """Frame, title, position ticks (-flank .. flank - 1 under positions 0 .. length - 1), dashed grid between positions (the cut sites),
    y ticks.  Returns the data -> page transform."""
    x0, y0, w, h = 70.0, 60.0, pg.w - 100.0, pg.h - 130.0
    pad = 0.05 * (ymax - ymin) if ymax > ymin else 1.0
    ymin, ymax = ymin - pad, ymax + pad
    sx = lambda x: x0 + w * (np.asarray(x, dtype=float) / max(1, length - 1))
    sy = lambda y: y0 + h * ((np.asarray(y, dtype=float) - ymin) / (ymax - ymin))
    cy = pg.h - 32.0
    for i, line in enumerate(title.split("\n")):
        pg.text(pg.w / 2.0, cy - 18.0 * i, line, size=14, bold=True, align="center")
    flank = int(length / 2.0)
    pg.color((0.8, 0.8, 0.8))
    pg.line_style(0.5, dash=(3, 3))
    for i in range(length - 1):
        pg.polyline([sx(i + 0.5)] * 2, [y0, y0 + h])
    pg.color((0, 0, 0))
    pg.line_style(0.8)
    pg.rect(x0, y0, w, h)
    for i in range(length):
        pg.polyline([sx(i)] * 2, [y0, y0 - 4.0])
        pg.text(float(sx(i)), y0 - 15.0, i - flank, size=7, align="center")
    for v in _nice_ticks(ymin, ymax):
        pg.polyline([x0, x0 - 4.0], [sy(v)] * 2)
        pg.text(x0 - 7.0, float(sy(v)) - 3.0, "%g" % round(v, 10), size=8, align="right")
    pg.text(x0 + w / 2.0, 22.0, xlabel, size=10, align="center")
    pg.text(20.0, y0 + h / 2.0, ylabel, size=10, align="center", rotate=True)
    return sx, sy, (x0, y0, w, h)


def _legend(pg, box, entries):
    x0, y0, w, _h = box
    lx, ly = x0 + w - 110.0, y0 + 10.0 + 11.0 * (len(entries) - 1)
    for label, rgb, width, dash in entries:
        pg.color(rgb)
        pg.line_style(width, dash)
        pg.polyline([lx, lx + 22.0], [ly + 3.0, ly + 3.0])
        pg.text(lx + 27.0, ly, label, size=8)
        ly -= 11.0


def pssm_page(matrix, title):
    """plot_pssm (atacorrect_functions.py:424-463): PSSM score of A, T, C, G per position; the cut site between flank - 1 and flank."""
    matrix = np.asarray(matrix, dtype=float)
    length = matrix.shape[1]
    vals = matrix[:4][np.isfinite(matrix[:4])]
    pg = _Page()
    sx, sy, box = _axes(pg, title, "Position from cutsite", "PSSM score", length, float(vals.min()) if vals.size else 0.0,
                        float(vals.max()) if vals.size else 1.0)
    xs = np.arange(length)
    for nuc in range(4):
        pg.color(COLORS[nuc])
        pg.line_style(1.5)
        pg.polyline(sx(xs), sy(matrix[nuc]))
    pg.color((0, 0, 0))
    pg.line_style(2.0)
    pg.polyline([sx(int(length / 2.0) - 0.5)] * 2, [box[1], box[1] + box[3]])
    _legend(pg, box, [(NAMES[n], COLORS[n], 1.5, None) for n in range(4)])
    return pg


def correction_page(pre_mat, post_mat, title):
    """plot_correction (:466-509): nucleotide frequencies before (dashed, thin) and after (solid) the correction."""
    pre_mat, post_mat = np.asarray(pre_mat, dtype=float), np.asarray(post_mat, dtype=float)
    length = pre_mat.shape[1]
    both = np.concatenate([pre_mat[:4].ravel(), post_mat[:4].ravel()])
    both = both[np.isfinite(both)]
    pg = _Page()
    sx, sy, box = _axes(pg, title, "Position from cutsite", "Nucleotide frequency", length, float(both.min()) if both.size else 0.0,
                        float(both.max()) if both.size else 1.0)
    xs = np.arange(length)
    for nuc in range(4):
        pg.color(tuple(0.5 + 0.5 * c for c in COLORS[nuc]))          # alpha 0.5 over white
        pg.line_style(1.0, dash=(4, 2))
        pg.polyline(sx(xs), sy(pre_mat[nuc]))
    for nuc in range(4):
        pg.color(COLORS[nuc])
        pg.line_style(2.0)
        pg.polyline(sx(xs), sy(post_mat[nuc]))
    _legend(pg, box, [(NAMES[n], COLORS[n], 2.0, None) for n in range(4)] +
            [("pre-correction", (0, 0, 0), 1.0, (4, 2)), ("post-correction", (0, 0, 0), 1.0, None)])
    return pg


def write_report(path, bias_obj, pre_bias, post_bias, strands=("forward", "reverse")):
    """The pages of the reference's figure PDF, in its order: bias motif per strand, then the pre / post comparison per strand."""
    pages = [pssm_page(bias_obj.bias[s].pssm, "Tn5 insertion bias of reads ({0})".format(s)) for s in strands]
    pages += [correction_page(pre_bias[s].bias_pwm, post_bias[s].bias_pwm, "Nucleotide frequencies in corrected reads\n({0} strand)".format(s))
              for s in strands]
    write_pdf(path, pages)
    return path
