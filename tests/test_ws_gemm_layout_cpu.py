"""CPU checks of the shared-memory addressing of the warp-specialised GEMM (csrc/dgemm_ws.cuh): with the
128-byte TMA swizzle and the k-permutation k(s, t4) = 8 (t4 >> 1) + (t4 & 1) + 2 ((t4 + s) & 3), every 8-byte fragment load
of either operand layout touches 16 distinct 8-byte bank pairs per half warp, each k of a 16-wide k-block is used exactly once,
and the per-lane offsets the kernel computes address exactly the element the DMMA fragment needs."""
import itertools


def kperm(s, t4):
    return 8 * (t4 >> 1) + (t4 & 1) + 2 * ((t4 + s) & 3)


def swz_offset_kc(row, k):
    """K-contiguous tile: 128-byte rows of 16 doubles; chunk (k >> 1) lands at (k >> 1) ^ (row & 7).  Offset in doubles."""
    return row * 16 + (((k >> 1) ^ (row & 7)) << 1) + (k & 1)


def swz_offset_box(k, col):
    """16 x 16 box, row = k (128 bytes), chunk (col >> 1) ^ (k & 7).  Offset in doubles inside the box."""
    return k * 16 + (((col >> 1) ^ (k & 7)) << 1) + (col & 1)


def test_every_k_of_a_block_is_used_exactly_once():
    ks = sorted(kperm(s, t4) for s in range(4) for t4 in range(4))
    assert ks == list(range(16))


def test_fragment_loads_are_bank_conflict_free_per_half_warp():
    for s in range(4):
        for half in range(2):
            lanes = [(g8, t4) for g8 in range(4 * half, 4 * half + 4) for t4 in range(4)]
            kc = {swz_offset_kc(g8, kperm(s, t4)) % 16 for g8, t4 in lanes}            # K-contiguous: row = g8, k permuted
            for parity in range(2):                                                    # fragment parity = which half of a box
                box = {swz_offset_box(kperm(s, t4), 8 * parity + g8) % 16 for g8, t4 in lanes}
                assert len(box) == 16, (s, half, parity)
            assert len(kc) == 16, (s, half)


def test_kernel_lane_offsets_address_the_fragment_elements():
    """a_off / b_off of the kernel (copied formulas) against the swizzle functions, for the NWM = 2 / NWN = 4 shapes."""
    for s, g8, t4, w in itertools.product(range(4), range(8), range(4), range(4)):
        kk = kperm(s, t4)
        kc = g8 * 16 + ((((kk >> 1) ^ g8)) << 1) + (kk & 1)
        mc = kk * 16 + ((((g8 >> 1) ^ (kk & 7))) << 1) + (g8 & 1)
        # K-contiguous operand, fragment f = i*NW + w  ->  row 8 f + g8
        for i in range(3):
            f = i * 4 + w
            assert kc + w * 128 + i * 4 * 128 == swz_offset_kc(8 * f + g8, kk)
            # boxed operand: fragment f sits in box f >> 1, columns 8 (f & 1) + g8
            off = (mc ^ ((w & 1) << 3)) + (w >> 1) * 256 + i * 2 * 256
            assert off == (f >> 1) * 256 + swz_offset_box(kk, 8 * (f & 1) + g8)


def test_bench_reads_its_roofline_inputs_from_tracked_profiles():
    import bench
    per_triple, src = bench.tracked_gemm_traffic()
    assert src and src.startswith("profiles/") and 1.0e9 < per_triple < 1.4e9   # 1.10 GB algorithmic, ~1.17 GB measured
    e_t, secs, o, v, rec = bench.committed_n1_e2e()
    assert (o, v) == (40, 400) and rec.startswith("profiles/") and 20.0 < secs < 30.0 and e_t == -22745304.824729037
    assert bench.make_config(40, 400, 8) == bench.make_config(40, 400, 8)      # both arms print this very dict
