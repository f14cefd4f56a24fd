"""Closed-form stage-2 (band -> bidiagonal) window schedule -- host-side mirror of the index
arithmetic in csrc/stage2_chase.cu (SURVEY 8a''; reference recurrence svd_parallel.h:652-688).

With c = band (= w-1 where w = band+1 is the reference's internal width, svd_parallel.h:648),
sweep i = 0..n-2 runs pairs k = -1, 0, 1, ...:

  RIGHT(i,k): rows [i+1+k*c, i+1+(k+2)*c)   cols [i+1+(k+1)*c, i+1+(k+2)*c)   (k=-1: rows [i, i+w))
  LEFT (i,k): rows [i+1+(k+1)*c, i+1+(k+2)*c) cols [i+1+(k+1)*c, i+1+(k+3)*c)

every bound clamped to n separately; a window with an empty column range is skipped.  The number of
chase pairs after the top pair is floor((n - min(i+2w-1, n)) / c) + 1 (integer division, as in the
reference, which is what leaves the last bulge un-chased -- SURVEY 0.3).
"""


def stage2_windows(n: int, band: int):
    """Yields (kind, i1, i2, j1, j2, sweep); kind 0 = right-applied (reflector from first row),
    1 = left-applied (reflector from first column)."""
    c = band
    w = band + 1
    for i in range(n - 1):
        yield (0, i, min(i + w, n), i + 1, min(i + w, n), i)
        yield (1, i + 1, min(i + w, n), i + 1, min(i + 2 * w - 1, n), i)
        if c < 1:
            continue
        npairs = (n - min(i + 2 * w - 1, n)) // c + 1
        for k in range(npairs):
            r0 = min(i + 1 + k * c, n)
            r1 = min(i + 1 + (k + 1) * c, n)
            r2 = min(i + 1 + (k + 2) * c, n)
            c3 = min(i + 1 + (k + 3) * c, n)
            if r2 > r1:                      # RIGHT: cols [r1, r2)
                yield (0, r0, r2, r1, r2, i)
            if c3 > r1:                      # LEFT: rows [r1, r2) cols [r1, c3)
                yield (1, r1, r2, r1, c3, i)


def stage2_pair_count(n: int, band: int, sweep: int) -> int:
    """Number of (RIGHT,LEFT) pairs of a sweep including the top pair."""
    w = band + 1
    return 1 + (n - min(sweep + 2 * w - 1, n)) // band + 1
