"""The barrier rules of k_bev_tma's consumer warps (cameracalibration_b200/csrc/bevk_bev_tma.cuh: D_SYNC / D_ROWS, the
producer's unit loop), checked on a model: which accumulator words which warp touches, and whether every pair of
conflicting accesses by two different warps is separated by a CTA barrier both of them pass.

Model (32 x 32 accumulator words per frame-set, eight consumer warps):
  * an item with the lanes along canvas x (orientation 0): warp w reads and writes rows w, w+8, w+16, w+24; along canvas y
    (orientation 1): columns w, w+8, w+16, w+24;
  * interior write-out: warp w reads rows w, w+8, w+16, w+24 (tile_out_row32); generic write-out (edge tiles, BALANCE):
    thread t -> row t / 8, so warp w reads rows 4w .. 4w+3;
  * barriers (bar.sync among the eight warps): before a slot with D_SYNC, and before the write-out unless the slot that
    ends the unit carries D_ROWS.  Rules as the producer applies them: D_SYNC on the first slot of a unit iff its first
    item has orientation 1 or the previous unit left through the generic write-out; D_SYNC whenever the orientation
    changes inside a unit; D_ROWS iff no item of the unit had orientation 1 and the tile takes the interior write-out.
Two accesses are ordered iff a barrier lies between them in program order (all warps pass the same barriers), i.e. iff
their barrier epochs differ.  The GPU counterpart is compute-sanitizer racecheck (profiles/r02_sanitizer_racecheck.log)."""
import itertools
import random

WARPS, T = 8, 32


def words_of_item(w, orient):
    lines = [w + 8 * k for k in range(4)]
    return {(y, x) for y in lines for x in range(T)} if orient == 0 else {(y, x) for x in lines for y in range(T)}


def words_of_writeout(w, interior):
    rows = [w + 8 * k for k in range(4)] if interior else [4 * w + k for k in range(4)]
    return {(y, x) for y in rows for x in range(T)}


def simulate(units, first_sync=lambda orient, prev_generic: orient == 1 or prev_generic,
             rows_only=lambda columns, interior: not columns and interior, none_resets=False):
    """units: list of (interior: bool, [orientations of the items]) -> list of races (epoch, word, warp a, warp b)."""
    epoch = 0
    last = {}            # word -> list of (epoch, warp, is_write) of the accesses since the word's last barrier-separated state
    races = []

    def access(words, warp, write):
        for wd in words:
            for (e, a, wr) in last.get(wd, ()):
                if a != warp and e == epoch and (wr or write):
                    races.append((epoch, wd, a, warp))
            last.setdefault(wd, []).append((epoch, warp, write))

    prev_generic = False
    for interior, orients in units:
        columns = False
        prev_orient = None
        for i, o in enumerate(orients):
            if (i == 0 and first_sync(o, prev_generic)) or (prev_orient is not None and prev_orient != o):
                epoch += 1
            prev_orient = o
            columns |= o == 1
            for w in range(WARPS):
                access(words_of_item(w, o), w, True)        # stores (first camera) or read-modify-write
        if not orients:                                      # tile without a camera: zeros
            if not interior:                                 # posted with D_SYNC, one more barrier before the generic write-out,
                epoch += 2                                   # which still loads the accumulator words it replaces by zeros
                for w in range(WARPS):
                    access(words_of_writeout(w, False), w, False)
                prev_generic = True
            elif none_resets:                                # the rule as first written: prev_generic = not interior
                prev_generic = False
            continue                                         # interior: no barrier, no access, what was unfenced stays unfenced
        if not rows_only(columns, interior):
            epoch += 1
        for w in range(WARPS):
            access(words_of_writeout(w, interior), w, False)
        prev_generic = not interior
        # forget what two barriers have separated for good (keeps the model small)
        for wd in list(last):
            last[wd] = [a for a in last[wd] if a[0] >= epoch - 1]
    return races


def all_units(max_items=3):
    for interior in (True, False):
        for n in range(max_items + 1):
            for orients in itertools.product((0, 1), repeat=n):
                yield (interior, list(orients))


def test_every_pair_triple_and_random_sequence_of_units_is_race_free():
    units = list(all_units())
    for a, b in itertools.product(units, repeat=2):
        assert not simulate([a, b]), (a, b)
    small = list(all_units(2))
    for a, b, c in itertools.product(small, repeat=3):
        assert not simulate([a, b, c]), (a, b, c)
    rng = random.Random(7)
    for _ in range(300):
        seq = [rng.choice(units) for _ in range(rng.randint(3, 8))]
        assert not simulate(seq), seq


def test_the_model_sees_a_missing_barrier():
    """Each rule is needed: without it the model reports a race (so the model is not vacuous)."""
    # first item along y right after a rows-only interior unit: other warps' columns cross rows still being written out
    assert simulate([(True, [0]), (True, [1])], first_sync=lambda o, g: False)
    # generic write-out reads rows 4w..4w+3: the next unit's first orientation-0 item (rows w+8k) needs the barrier
    assert simulate([(False, [0]), (True, [0])], first_sync=lambda o, g: o == 1)
    # a unit with a column item must meet before its row-wise write-out
    assert simulate([(True, [0, 1])], rows_only=lambda c, i: i)
    # and a rows-only unit on an edge tile (generic write-out) as well
    assert simulate([(False, [0])], rows_only=lambda c, i: not c)
    # an empty interior tile must not clear the "generic write-out unfenced" state (the rule as first written did:
    # edge tile -> tile without a camera -> tile whose first item runs along x)
    assert simulate([(False, [0]), (True, []), (True, [0])], none_resets=True)
    assert not simulate([(False, [0]), (True, []), (True, [0])])
