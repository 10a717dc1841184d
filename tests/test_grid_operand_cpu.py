"""CPU tests of two host-side planners of round 2 (no GPU):

* ``psm_grid_operand_plan`` -- the TMA box plan that replaces the block extraction of SMC:464-492 / GRAD:479-516 (the PCA
  projection fetches its A tiles from the grid planes): every block exactly once, from its own origin, in as many 128-row tiles
  as the extracted operand needs;
* ``psm_send_map_build`` -- the ghost-cell send map of a sharded handle (replaces the gather to rank 0 of PMP:258): the lookup the
  prep kernels do (bitmap word + popcount + entry chain), restated in NumPy, returns exactly the (peer, slot) pairs of the lists.
"""
import ctypes as C

import numpy as np
import pytest

import psm_b200
from psm_b200 import _capi as capi
from psm_b200.surrogate import compile_plan


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def grid_plan(origins, stride):
    lib = capi.load()
    by0 = np.ascontiguousarray(origins[:, 0], np.int32)
    bx0 = np.ascontiguousarray(origins[:, 1], np.int32)
    B = by0.size
    tiles, gx, gy, ns = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
    row_src = np.full(B, -7, np.int32)
    segs = np.zeros((4 * B + 8, 5), np.int32)
    rc = lib.psm_grid_operand_plan(B, _p(by0, C.c_int32), _p(bx0, C.c_int32), stride, C.byref(tiles), C.byref(gx), C.byref(gy),
                                   C.byref(ns), _p(row_src, C.c_int32), _p(segs, C.c_int32), segs.shape[0])
    assert rc == 0
    return tiles.value, gx.value, gy.value, row_src, segs[:ns.value]


CASES = [   # (variant, H, W, overlap): BASELINE configs[0..2], c5, the thesis module's plan, a non-square grid
    ('deltaU_to_deltaP', 400, 3000, 32), ('deltaU_to_deltaP', 1000, 1000, 32), ('U_to_gradP', 1000, 1000, 96),
    ('deltaU_to_deltaP', 2000, 2000, 32), ('thesis', 1000, 1000, 12), ('U_to_gradP', 520, 776, 96), ('deltaU_to_deltaP', 500, 420, 32),
]


@pytest.mark.parametrize('variant,H,W,overlap', CASES)
def test_box_plan_covers_every_block_once(variant, H, W, overlap):
    plan = compile_plan(variant, H, W, np.ones((H, W), np.uint8), 128, overlap)
    org = plan['origins']
    B, st = org.shape[0], 128 - overlap
    tiles, gx, gy, row_src, segs = grid_plan(org, st)
    assert tiles >= 1 and gx >= 4
    # as many tiles as the extracted operand needs (padding never costs a whole tile on the plans of the reference)
    assert tiles == -(-B // 128)
    # block -> operand row is injective and inside the tiles
    assert np.unique(row_src).size == B and row_src.min() >= 0 and row_src.max() < tiles * 128
    # every segment fetches the blocks it claims to: rebuild block <- row from the segments alone
    row_to_block = {int(r): b for b, r in enumerate(row_src)}
    seen = np.zeros(B, np.int32)
    for t, kind, row, x, y in segs:
        n = {0: gx, 1: gy, 2: 1}[int(kind)]
        for i in range(n):
            b = row_to_block[int(t) * 128 + int(row) + i]
            assert org[b, 1] == x + (i * st if kind == 0 else 0)
            assert org[b, 0] == y + (i * st if kind == 1 else 0)
            seen[b] += 1
        assert row + n <= 128
    assert (seen == 1).all()
    # the boxes carry most of the operand: a k-block of a tile is a handful of TMA instructions, not 128
    per_tile = np.bincount(segs[:, 0], minlength=tiles)
    assert per_tile.max() <= 32, 'one box per lane of the producer warp'


def test_box_plan_declines_scattered_blocks():
    rng = np.random.default_rng(0)
    org = np.stack([rng.integers(0, 800, 9) * 4, rng.integers(0, 800, 9) * 4 + 1], 1).astype(np.int32)
    tiles, gx, gy, row_src, segs = grid_plan(org, 96)
    assert tiles == 0 and segs.shape[0] == 0            # no run of >= 4 blocks: the handle keeps the extracted operand


def send_map(n_cells, lists):
    lib = capi.load()
    world = len(lists)
    ptr = np.zeros(world + 1, np.int64)
    ptr[1:] = np.cumsum([len(x) for x in lists])
    idx = np.ascontiguousarray(np.concatenate([np.asarray(x, np.int32) for x in lists]) if ptr[-1] else np.zeros(0, np.int32))
    words = np.zeros(((n_cells + 31) // 32 + 1, 2), np.uint32)
    entries = np.zeros((max(int(ptr[-1]), 1), 2), np.int32)
    ne = C.c_int64()
    rc = lib.psm_send_map_build(n_cells, world, _p(ptr, C.c_int64), _p(idx, C.c_int32) if idx.size else None,
                                _p(words, C.c_uint32), _p(entries, C.c_int32), C.byref(ne))
    assert rc == 0 and ne.value == ptr[-1]
    return words, entries


def lookup(words, entries, i):
    """What prep_push_cell does for owned cell i (csrc/psm_kernels.cu)."""
    bits, prefix = int(words[i >> 5, 0]), int(words[i >> 5, 1])
    bit = 1 << (i & 31)
    if not bits & bit:
        return []
    k = prefix + bin(bits & (bit - 1)).count('1')
    out = []
    while True:
        e0, e1 = int(entries[k, 0]), int(entries[k, 1])
        out.append((e0 & 0xFF, e1))
        if not e0 & 0x100:
            return out
        k = e0 >> 9


@pytest.mark.parametrize('n_cells,world,seed', [(1000, 2, 0), (4097, 4, 1), (70001, 8, 2), (31, 3, 3)])
def test_send_map_lookup_equals_the_lists(n_cells, world, seed):
    rng = np.random.default_rng(seed)
    lists = []
    for p in range(world):
        k = int(rng.integers(0, max(2, n_cells // 3)))
        # sorted unique per peer (as the partitioner delivers them), overlapping between peers: some cells go to several ranks
        lists.append(np.sort(rng.choice(n_cells, size=min(k, n_cells), replace=False)))
    lists[0] = np.union1d(lists[0], [0, n_cells - 1]).astype(np.int64)        # first and last cell, last word partially filled
    words, entries = send_map(n_cells, lists)
    expect = {}
    for p, lst in enumerate(lists):
        for slot, c in enumerate(lst):
            expect.setdefault(int(c), []).append((p, slot))
    multi = sum(1 for v in expect.values() if len(v) > 1)
    assert world < 3 or multi > 0, 'the case must exercise the entry chain'
    for i in range(n_cells):
        assert sorted(lookup(words, entries, i)) == sorted(expect.get(i, [])), i


def test_send_map_empty_and_bad_input():
    words, entries = send_map(100, [[], []])
    assert not words[:, 0].any()
    lib = capi.load()
    ptr = np.array([0, 1], np.int64)
    bad = np.array([100], np.int32)
    ne = C.c_int64()
    assert lib.psm_send_map_build(100, 1, _p(ptr, C.c_int64), _p(bad, C.c_int32), None, None, C.byref(ne)) < 0     # cell id out of range


@pytest.mark.parametrize('world', [2, 3])
def test_send_map_push_fills_every_ghost_region(world):
    """The fused flow's first exchange on a real partition, in NumPy: every rank converts its cells and pushes, cell by cell
    through the send map, into the consumers' ghost regions (slot = position in the consumer's ghost list from that owner) --
    afterwards each rank's [owned | ghost] field equals the global field at its local-to-global ids (what NCCL send/recv or
    the reference's gather to rank 0, PMP:258, would have delivered)."""
    from psm_b200 import synthetic as syn, tables as ptables, shard as pshard
    mesh = syn.make_mesh(seed=5, H=500, W=420, nx=160, ny=200, R=0.12)
    F = syn.make_fields(mesh, seed=5)
    t = ptables.build_tables(mesh['cells'], mesh['top'], mesh['obst'], F['p_prev'])
    shards = pshard.partition(t, mesh['cells'], world)
    field = np.stack([F['dUx'], F['dUy']], 1)                       # the global cell field
    uv = [np.full((s['n_owned'] + s['n_ghost'], 2), np.nan) for s in shards]
    for s in shards:
        uv[s['rank']][:s['n_owned']] = field[s['owned_ids']]
    pushed = 0
    for s in shards:
        r = s['rank']
        words, entries = send_map(s['n_owned'], [s['cell_send_idx'][s['cell_send_ptr'][p]:s['cell_send_ptr'][p + 1]] for p in range(world)])
        marked = np.flatnonzero([(int(words[i >> 5, 0]) >> (i & 31)) & 1 for i in range(s['n_owned'])])
        for i in marked:
            for peer, slot in lookup(words, entries, int(i)):
                sp = shards[peer]
                uv[peer][sp['n_owned'] + sp['cell_recv_ptr'][r] + slot] = uv[r][i]      # P2PArgs::uv_ghost[peer] + slot
                pushed += 1
    assert pushed == sum(s['n_ghost'] for s in shards) > 0
    for s in shards:
        l2g = np.concatenate([s['owned_ids'], s['ghost_ids']])
        assert np.array_equal(uv[s['rank']], field[l2g])
