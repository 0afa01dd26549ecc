"""CIELAB ab-gamut quantisation table (313 bins of 10x10 in the ab plane).

Replaces `utils/cielab.py:5-64` (`ABGamut`, `CIELAB.q_to_ab`) of the reference, which loads
`utils/gamut_pts.npy` through a cwd-relative path.  The 313 in-gamut bins of Zhang et al. (ECCV'16)
form 20 contiguous b-intervals, one per a-value, so the table is stored here as run-lengths and
expanded at import time; `q_to_ab[q]` is the bin centre, ordered by (a, b) exactly as
`ab[ab_gamut_mask] + AB_BINSIZE/2` orders it (`utils/cielab.py:62-64`).
"""
import numpy as np

AB_BINSIZE = 10
N_BINS = 313

# (a, b_lo, b_hi) -- inclusive, step 10
_GAMUT_ROWS = (
    (-90, 50, 90), (-80, 20, 90), (-70, 0, 90), (-60, -20, 90), (-50, -30, 100),
    (-40, -40, 100), (-30, -50, 100), (-20, -50, 100), (-10, -60, 100), (0, -70, 100),
    (10, -80, 90), (20, -80, 90), (30, -90, 90), (40, -100, 90), (50, -100, 80),
    (60, -110, 80), (70, -110, 80), (80, -110, 70), (90, -110, 70), (100, -90, 0),
)


def q_to_ab():
    """(313, 2) float32 bin centres in Lab units (not divided by 110)."""
    pts = [(a, b) for a, lo, hi in _GAMUT_ROWS for b in range(lo, hi + AB_BINSIZE, AB_BINSIZE)]
    out = np.asarray(pts, dtype=np.float32)
    assert out.shape == (N_BINS, 2)
    return out


Q_TO_AB = q_to_ab()
