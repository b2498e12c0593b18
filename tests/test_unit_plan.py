"""CPU tests of the feature-pair unit plan (mixdir_b200/csrc/engine.cu: make_units) through the
host-only C-ABI door mixdir_b200_unit_plan: the packed codes, the unit table T2 and the
marginalisation of the unit statistics must reproduce the per-feature gather
(src/mixdir_impl.cpp:77-82) and the per-feature M-step sums (R/mixdir_vi.R:85-91) exactly."""
import ctypes as C

import numpy as np
import pytest

from mixdir_b200 import _capi


def unit_plan(ncat, m, has_na=None):
    L = _capi.lib()
    ncat = np.ascontiguousarray(ncat, dtype=np.int32)
    J = ncat.size
    U, NR2 = C.c_int(), C.c_int()
    ip = C.POINTER(C.c_int)
    na = None if has_na is None else np.ascontiguousarray(has_na, dtype=np.uint8)
    nap = None if na is None else na.ctypes.data
    st = L.mixdir_b200_unit_plan(J, ncat.ctypes.data_as(ip), m, nap, C.byref(U), C.byref(NR2), None, None,
                                 None, None)
    if st != 0:
        return None
    NR = int((ncat + 1).sum())
    desc = np.zeros((U.value, 4), dtype=np.int32)
    gbase = np.zeros(U.value, dtype=np.int32)
    t2src = np.full((NR2.value, 2), -1, dtype=np.int32)
    marg = np.zeros((NR, 4), dtype=np.int32)
    st = L.mixdir_b200_unit_plan(J, ncat.ctypes.data_as(ip), m, nap, C.byref(U), C.byref(NR2),
                                 desc.ctypes.data, gbase.ctypes.data, t2src.ctypes.data, marg.ctypes.data)
    assert st == 0
    return dict(U=U.value, NR2=NR2.value, desc=desc, gbase=gbase, t2src=t2src, marg=marg)


def pack(codes, desc):
    out = np.zeros((codes.shape[0], desc.shape[0]), dtype=np.int64)
    for e, (oa, ob, mul, flags) in enumerate(desc):
        if oa < 0:
            continue
        out[:, e] = codes[:, oa] - (flags & 1)
        if ob >= 0:
            out[:, e] = out[:, e] * mul + (codes[:, ob] - ((flags >> 1) & 1))
    return out


@pytest.mark.parametrize("na_mode", ["all", "none", "some"])
@pytest.mark.parametrize("ncat,m", [
    ([16, 12, 8, 4, 2, 10] * 3, 9), ([16, 12, 8, 4, 2, 10] * 3, 4), ([2, 3, 4, 5, 6], 2),
    ([10] * 7, 3), ([1, 1, 1], 1), ([15, 15, 3], 1), ([120, 1, 1, 100], 2)])
def test_unit_plan_reproduces_gather_and_scatter(ncat, m, na_mode):
    rng = np.random.default_rng(7)
    ncat = np.asarray(ncat)
    J, N, K = ncat.size, 400, 3
    # features without any NA lose their NA row (has_na[j] == 0); "all" = the default plan
    has_na = {"all": None, "none": np.zeros(J, dtype=np.uint8),
              "some": (np.arange(J) % 3 == 0).astype(np.uint8)}[na_mode]
    up = unit_plan(ncat, m, has_na)
    assert up is not None
    rowoff = np.concatenate([[0], np.cumsum(ncat + 1)[:-1]])
    NR = int((ncat + 1).sum())
    lo = np.zeros(J, dtype=int) if has_na is None else 1 - has_na.astype(int)
    codes = np.stack([rng.integers(lo[j], c + 1, size=N) for j, c in enumerate(ncat)], axis=1)  # 0 = NA
    u = pack(codes, up["desc"])
    assert up["U"] % 4 == 0 and u.max() <= 255
    n_real = int((up["desc"][:, 0] >= 0).sum())
    assert n_real == J - m and int((up["desc"][:, 1] >= 0).sum()) == m
    # every feature appears exactly once
    used = sorted([int(d[0]) for d in up["desc"] if d[0] >= 0] + [int(d[1]) for d in up["desc"] if d[1] >= 0])
    assert used == list(range(J))
    # gather: sum over elements of T2[gbase + u] == sum over features of T[rowoff + x]
    T = rng.normal(size=(NR, K))
    T[rowoff] = 0.0  # NA rows
    src = up["t2src"]
    T2 = T[src[:, 0]] + np.where(src[:, 1:2] >= 0, T[np.maximum(src[:, 1], 0)], 0.0)
    g2 = up["gbase"][None, :] + u
    real = up["desc"][:, 0] >= 0
    assert g2.max() < up["NR2"]
    lhs = T2[g2[:, real]].sum(axis=1)
    rhs = T[rowoff[None, :] + codes].sum(axis=1)
    np.testing.assert_allclose(lhs, rhs, rtol=0, atol=1e-12)
    # padding elements point at a row whose T2 is zero (plans that keep every NA row: that is what
    # k_em_fused relies on; the unit kernel gives padding its own zero row instead)
    if has_na is None:
        assert np.all(T2[g2[:, ~real]] == 0.0)
    # scatter: unit statistics marginalised by marg == feature statistics
    z = rng.random(size=(N, K))
    S2 = np.zeros((up["NR2"], K))
    for e in np.flatnonzero(real):
        np.add.at(S2, g2[:, e], z)
    S = np.zeros((NR, K))
    for j in range(J):
        np.add.at(S, rowoff[j] + codes[:, j], z)
    for g in range(NR):
        start, stride, count, _ = up["marg"][g]
        if g in set(rowoff.tolist()):
            assert count == 0  # NA rows are never read (na.rm, R/mixdir_vi.R:88)
            continue
        got = S2[start + stride * np.arange(count)].sum(axis=0)
        np.testing.assert_allclose(got, S[g], rtol=1e-12, atol=1e-12)


def test_unit_plan_rejects_pairs_that_do_not_fit_a_byte():
    assert unit_plan([16, 16], 1) is None      # 17 * 17 = 289 codes
    assert unit_plan([15, 15], 1) is not None  # 16 * 16 = 256
    assert unit_plan([3, 3, 3], 2) is None     # more pairs than features allow
    ident = unit_plan([3, 4, 5], 0)
    assert ident["NR2"] == 15 and ident["U"] == 4  # identity: padded features, feature tables
    lean = unit_plan([3, 4, 5], 1, np.zeros(3, dtype=np.uint8))
    assert lean["NR2"] == 3 * 4 + 5  # (3 x 4) pair without NA rows + the single feature with 5


def unit_layout(ncat, m, KC):
    L = _capi.lib()
    ncat = np.ascontiguousarray(ncat, dtype=np.int32)
    J = ncat.size
    ip = C.POINTER(C.c_int)
    U, NR2, SR = C.c_int(), C.c_int(), C.c_int()
    assert L.mixdir_b200_unit_layout(J, ncat.ctypes.data_as(ip), m, None, KC, C.byref(U), C.byref(NR2),
                                     C.byref(SR), None, None, None, None) == 0
    desc = np.zeros((U.value, 4), dtype=np.int32)
    gbase = np.zeros(U.value, dtype=np.int32)
    slot = np.zeros(NR2.value, dtype=np.int32)
    eoff = np.zeros((4, U.value), dtype=np.int32)
    assert L.mixdir_b200_unit_layout(J, ncat.ctypes.data_as(ip), m, None, KC, C.byref(U), C.byref(NR2),
                                     C.byref(SR), desc.ctypes.data, gbase.ctypes.data,
                                     slot.ctypes.data, eoff.ctypes.data) == 0
    return dict(U=U.value, NR2=NR2.value, srows=SR.value, desc=desc, gbase=gbase, slot=slot, eoff=eoff)


@pytest.mark.parametrize("KC", [8, 16, 32, 4])
@pytest.mark.parametrize("ncat,m", [([16, 12, 8, 4, 2, 10] * 4, 10), ([16, 12, 8, 4, 2, 10] * 4, 0),
                                    ([2, 3, 4, 5, 6], 2), ([7], 0)])
def test_unit_layout_bank_positions_and_gather(ncat, m, KC):
    """The shared-memory placement of k_em_units (engine.cu: layout_units), emulated in numpy: every
    unit-table row has its own slot, element e sits at bank position e mod Q, the slots of one warp
    instruction cover the positions evenly under the slot rotation, and gathering through
    (eoff_rot, code) from the emulated shared-memory table returns T2[gbase + code]."""
    ncat = np.asarray(ncat)
    lay = unit_layout(ncat, m, KC)
    up = unit_plan(ncat, m) if m > 0 else None
    Q = 1 if KC >= 16 else 16 // KC
    RS = max(128, KC * 8)
    U, NR2, srows = lay["U"], lay["NR2"], lay["srows"]
    assert U % 4 == 0 and (up is None or NR2 == up["NR2"])
    nrows = np.zeros(U, dtype=int)  # rows per element, from its descriptor
    for e, (oa, ob, mul, _) in enumerate(lay["desc"]):
        if oa >= 0:
            nrows[e] = (ncat[oa] + 1) * (mul if ob >= 0 else 1)
    assert nrows.sum() == NR2
    slot = lay["slot"]
    assert len(set(slot.tolist())) == NR2 and slot.min() >= Q and slot.max() < srows * Q  # row 0 is reserved
    for e in range(U):
        if nrows[e]:
            sl = slot[lay["gbase"][e]:lay["gbase"][e] + nrows[e]]
            assert np.all(sl % Q == e % Q) and np.all(np.diff(sl) == Q)  # one position, consecutive rows
            assert lay["eoff"][0, e] == sl[0] * KC * 8
        else:
            assert lay["eoff"][0, e] == (e % Q) * KC * 8  # padding -> zero row of its position
    for r in range(4):
        for e in range(U):
            assert lay["eoff"][r, e] == lay["eoff"][0, (e & ~3) | ((e + r) & 3)]
    # balanced stacks: the tallest position stack is the table height
    heights = [1 + sum(nrows[e] for e in range(U) if e % Q == q) for q in range(Q)]
    assert srows == max(heights)
    # slots s = 0 .. 32/KC - 1 of one instruction at sub-step f: positions covered evenly
    RPW = 32 // KC
    for f in range(4):
        pos = [((0 & ~3) | ((f + (s & 3)) & 3)) % Q for s in range(RPW)]
        assert all(pos.count(q) == RPW // Q for q in range(Q))
    # emulated gather for class kc = 0: shared-memory table as doubles, slot * KC + kc
    rng = np.random.default_rng(1)
    T2 = rng.normal(size=NR2)
    Ts = np.zeros(srows * (RS // 8))
    Ts[slot * KC] = T2
    for e in range(U):
        for code in range(nrows[e]):
            addr = lay["eoff"][0, e] + code * RS  # bytes
            assert Ts[addr // 8] == T2[lay["gbase"][e] + code]
