"""Index maps of the register-resident kernels, restated in NumPy (CPU only).

The three-pass kernel (tike_b200/csrc/rpie_p3.cu) and the large-detector kernels
(large_k2r.cu, large_k13r.cu) rely on hand-derived ownership maps: which tile
entries a thread touches in each pass, which frequency a far-field slot holds,
where the measured pixel of that frequency was parked.  These tests recompute
the maps from the formulas in the kernels and check that they are what the
comments claim: bijections, free of shared-memory bank conflicts, and
consistent between writer and reader.  (The butterflies themselves are
checked in tests/csrc/fft_host_test.cu, the kernels on the GPU.)
"""
import numpy as np

ND, NT, P = 128, 512, 129


def _conflict_free_64bit(word_addr):
    """8-byte shared-memory accesses of one warp instruction: served per half
    warp; conflict free iff the 16 lanes of each half hit 16 different bank
    pairs (or the same address)."""
    word_addr = np.asarray(word_addr).reshape(-1, 16)
    for half in word_addr:
        pairs = (half // 2) % 16
        uniq = {}
        for a, p in zip(half, pairs):
            if p in uniq and uniq[p] != a:
                return False
            uniq[p] = a
    return True


def _conflict_free_32bit(word_addr):
    word_addr = np.asarray(word_addr)
    banks = word_addr % 32
    seen = {}
    for a, b in zip(word_addr, banks):
        if b in seen and seen[b] != a:
            return False
        seen[b] = a
    return True


def test_three_pass_ownerships_cover_the_tile_without_bank_conflicts():
    tid = np.arange(NT)
    lane, wu = tid & 31, tid >> 5
    # pass 1: rows wu + 16 k, columns lane + 32 a
    cover = np.zeros((ND, ND), int)
    for k in range(8):
        for a in range(4):
            r, c = wu + 16 * k, lane + 32 * a
            cover[r, c] += 1
            for w in range(16):
                sel = wu == w
                assert _conflict_free_64bit(2 * (r[sel] * P + c[sel]))
    assert (cover == 1).all()
    # pass 2: rows 16 (tid >> 6) + n, columns 32 ((tid >> 4) & 3) + 16 b + (tid & 15)
    cover[:] = 0
    for n in range(16):
        for b in range(2):
            r = 16 * (tid >> 6) + n
            c = 32 * ((tid >> 4) & 3) + 16 * b + (tid & 15)
            cover[r, c] += 1
            for w in range(16):
                sel = wu == w
                assert _conflict_free_64bit(2 * (r[sel] * P + c[sel]))
    assert (cover == 1).all()
    # pass 3: row 32 (wu & 3) + lane, columns 32 (wu >> 2) + 16 q + p
    cover[:] = 0
    for q in range(2):
        for p in range(16):
            r = 32 * (wu & 3) + lane
            c = 32 * (wu >> 2) + 16 * q + p
            cover[r, c] += 1
            for w in range(16):
                sel = wu == w
                assert _conflict_free_64bit(2 * (r[sel] * P + c[sel]))
    assert (cover == 1).all()


def test_pattern_staging_swizzle_matches_the_far_field_owners():
    """The measured pattern is written in natural order (pixel = tid + 512 j) to
    D[fr * 128 + (fc ^ sw(fr))], sw(fr) = ((fr >> 3) & 15) | ((fr & 1) << 4); the
    pass-3 owner of slot (q, p1) reads D[fr_own * 128 + (fc_own ^ lane)] with
    fr_own = 8 (lane & 15) + 2 (wu & 3) + (lane >> 4), fc_own = (wu >> 2) + 4 q +
    8 p1.  Both sides conflict free, every pixel read exactly once, and the
    pixel read is the frequency the slot holds (row slot r <-> (r >> 4) + 8 (r &
    15), column slot 32 a1 + 16 b1 + p1 <-> a1 + 4 b1 + 8 p1)."""
    tid = np.arange(NT)
    lane, wu = tid & 31, tid >> 5
    D = np.full(ND * ND, -1, int)  # which pixel sits at each float slot
    for j in range(32):
        fr = (tid >> 7) + 4 * j
        fc = tid & (ND - 1)
        sw = ((fr >> 3) & 15) | ((fr & 1) << 4)
        addr = fr * ND + (fc ^ sw)
        assert (D[addr] == -1).all()
        D[addr] = fr * ND + fc
        for w in range(16):
            assert _conflict_free_32bit(addr[wu == w])
    assert (D >= 0).all()
    seen = np.zeros(ND * ND, int)
    for q in range(2):
        for p1 in range(16):
            fr_own = 8 * (lane & 15) + 2 * (wu & 3) + (lane >> 4)
            fc_own = (wu >> 2) + 4 * q + 8 * p1
            addr = fr_own * ND + (fc_own ^ lane)
            for w in range(16):
                assert _conflict_free_32bit(addr[wu == w])
            pix = D[addr]
            # the slot this thread holds after pass 3
            r = 32 * (wu & 3) + lane
            c = 32 * (wu >> 2) + 16 * q + p1
            f_row = (r >> 4) + 8 * (r & 15)
            f_col = (c >> 5) + 4 * ((c >> 4) & 1) + 8 * (c & 15)
            assert (pix == f_row * ND + f_col).all()
            seen[pix] += 1
    assert (seen == 1).all()


def test_large_k2_tiles_are_conflict_free():
    """large_k2r.cu: 8 rows per CTA, 128 threads (r = tid >> 4, h = tid & 15).
    256: element c of a row at c + c / 16 (pitch 272); stage 1 touches h + 16 k,
    stage 2 touches 16 h + n.  512: c + c / 32 (pitch 528); h + 16 q + 32 k and
    32 h + n."""
    tid = np.arange(128)
    r, h = tid >> 4, tid & 15
    for nd, pitch, shift, r2 in ((256, 272, 4, 16), (512, 528, 5, 32)):
        def idx(row, c):
            return row * pitch + c + (c >> shift)
        cover = np.zeros((8, nd), int)
        for k in range(16):
            for q in range(r2 // 16):
                c = h + 16 * q + r2 * k
                cover[r, c] += 1
                for w in range(4):
                    sel = (tid >> 5) == w
                    assert _conflict_free_64bit(2 * idx(r[sel], c[sel]))
        assert (cover == 1).all()
        cover[:] = 0
        for n in range(r2):
            c = r2 * h + n
            cover[r, c] += 1
            for w in range(4):
                sel = (tid >> 5) == w
                assert _conflict_free_64bit(2 * idx(r[sel], c[sel]))
        assert (cover == 1).all()
        assert idx(7, nd - 1) < 8 * pitch
