"""CPU checks of the work-unit bookkeeping of the pair stage (opensph_b200/csrc/pair_tiled.cu), restated in Python:

* k_units cuts the column-ordered targets of a double row into units of <= 128 (a column may be split, the x-range of
  a unit is bounded by TILE_X): every target must belong to exactly one unit and a unit must need exactly the columns
  its descriptor names;
* the unit order interleaves UNIT_KBLOCK double rows in z per cy: the map p -> (k, cy) must be a bijection.

The CUDA code itself is exercised by the GPU parity tests; these tests pin the algorithm the kernels implement, so that
a change of the packing rules is a conscious one."""
import random

TILE_T, TILE_X, UNIT_KBLOCK = 128, 20, 4


def walk(col):
    """k_units: lane 0's walk over the column counts of one double row -> [(cA, skip, span, targets)]."""
    units = []
    c_a = c_last = 0
    skip_a = taken = 0
    for c, cnt in enumerate(col):
        avail, skip = cnt, 0
        while avail > 0:
            if taken > 0 and c - c_a + 3 > TILE_X:
                units.append((c_a, skip_a, c_last - c_a, taken))
                taken = 0
            if taken == 0:
                c_a, skip_a = c, skip
            take = min(avail, TILE_T - taken)
            taken += take
            skip += take
            avail -= take
            c_last = c
            if taken == TILE_T:
                units.append((c_a, skip_a, c_last - c_a, taken))
                taken = 0
    if taken > 0:
        units.append((c_a, skip_a, c_last - c_a, taken))
    return units


def check_partition(col):
    units = walk(col)
    seq = [(c, i) for c in range(len(col)) for i in range(col[c])]
    got = []
    for c_a, skip_a, span, taken in units:
        assert 0 < taken <= TILE_T and span <= TILE_X - 3
        s = [(c, i) for c in range(c_a, c_a + span + 1) for i in range(col[c])][skip_a:skip_a + taken]
        assert len(s) == taken
        assert s[0][0] == c_a and s[-1][0] == c_a + span  # the descriptor names exactly the columns in use
        got += s
    assert got == seq


def test_units_partition_the_targets_of_a_double_row():
    random.seed(1)
    for _ in range(1500):
        n = random.randint(1, 300)
        mode = random.random()
        if mode < 0.3:
            col = [random.choice([0, 0, 1, 5, 25, 30, 130, 300, 1000]) for _ in range(n)]
        elif mode < 0.6:
            col = [random.randint(0, 3) for _ in range(n)]  # sparse rim: the x-range limit cuts the units
        else:
            col = [random.randint(20, 30) for _ in range(n)]  # the lattice interior
        check_partition(col)


def test_units_are_full_in_the_lattice_interior():
    col = [25] * 200
    units = walk(col)
    assert all(t == TILE_T for (_, _, _, t) in units[:-1])
    assert len(units) == -(-sum(col) // TILE_T)


def test_unit_order_is_a_bijection_of_the_double_rows():
    for dimy in (1, 2, 5, 13):
        for nk in (1, 2, 3, 4, 5, 7, 8, 9, 16, 17):
            seen = set()
            for p in range(nk * dimy):
                kb = min(p // (UNIT_KBLOCK * dimy), (nk - 1) // UNIT_KBLOCK)
                h_b = min(UNIT_KBLOCK, nk - kb * UNIT_KBLOCK)
                rem = p - kb * UNIT_KBLOCK * dimy
                cy, k = rem // h_b, kb * UNIT_KBLOCK + rem % h_b
                assert 0 <= cy < dimy and 0 <= k < nk
                seen.add((k, cy))
            assert len(seen) == nk * dimy
