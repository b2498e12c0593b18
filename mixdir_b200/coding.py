"""Input coding of mixdir() (reference: /root/reference/R/mixdir.R:91-118), host side.

Columns of strings (None / NaN = NA) become category codes 1..C_j with 0 for NA;
`categories[j]` are the levels: the sorted distinct observed strings of the rows
actually passed (R: as.factor(as.character(col)), R/mixdir.R:99-102).  With
na_handle="category" every column gets the extra level "(Missing)" and NA maps
to it (R/mixdir.R:103-106).
"""
from __future__ import annotations

import numpy as np

MISSING_LEVEL = "(Missing)"


def _is_na(v) -> bool:
    return v is None or (isinstance(v, float) and v != v)


def code_columns(columns, na_handle="ignore", levels=None):
    """columns: sequence of J sequences (length N) -> (codes[N,J] uint8|uint16, categories).

    levels: optional explicit level lists (R factor input keeps its own levels,
    including unused ones, R/mixdir.R:100).

    Level ORDER when levels are derived here: Python's sorted() over str(v).  R derives them with
    as.factor(as.character(col)), i.e. R's number formatting (1.0 -> "1", here "1.0") and the
    session locale's collation (mixed case orders differently from code-point order).  The order
    only permutes the category axis of phi / category_prob; a caller that needs R's exact order
    (to compare category_prob position by position) passes `levels` explicitly, which is what the
    parity tests do.
    """
    if na_handle not in ("ignore", "category"):
        raise ValueError("na_handle must be 'ignore' or 'category'")  # match.arg, R/mixdir.R:91
    J = len(columns)
    N = len(columns[0]) if J else 0
    cats, coded = [], []
    for j, col in enumerate(columns):
        vals = [None if _is_na(v) else str(v) for v in col]
        lev = list(levels[j]) if levels is not None and levels[j] is not None else \
            sorted({v for v in vals if v is not None})
        if na_handle == "category":
            lev = lev + [MISSING_LEVEL]
            vals = [MISSING_LEVEL if v is None else v for v in vals]
        if len(lev) == 0:
            # R/mixdir.R:112-115
            raise ValueError(f"Column {j} is empty (i.e. all values are NA). Please fix this.")
        idx = {s: i + 1 for i, s in enumerate(lev)}
        coded.append([idx.get(v, 0) if v is not None else 0 for v in vals])
        cats.append(lev)
    cmax = max((len(c) for c in cats), default=0)
    dt = np.uint8 if cmax <= 255 else np.uint16
    codes = np.zeros((N, J), dtype=dt)
    for j in range(J):
        codes[:, j] = coded[j]
    return codes, cats


def recode_for_predict(columns_by_name, names, categories, na_handle="ignore"):
    """R/predict.R:64-84: code new data against the TRAINING categories by column name.

    Missing column -> all NA; unseen level -> NA (+ it is reported back so the
    caller can warn like R/predict.R:75-80).
    """
    n = None
    for v in columns_by_name.values():
        n = len(v)
        break
    n = n or 1
    cmax = max(len(c) for c in categories)
    codes = np.zeros((n, len(names)), dtype=np.uint8 if cmax <= 255 else np.uint16)
    unseen = {}
    for j, name in enumerate(names):
        col = columns_by_name.get(name)
        if col is None:
            continue
        idx = {s: i + 1 for i, s in enumerate(categories[j])}
        for i, v in enumerate(col):
            if _is_na(v):
                v = MISSING_LEVEL if na_handle == "category" else None
            if v is None:
                continue
            c = idx.get(str(v), 0)
            if c == 0:
                unseen.setdefault(name, set()).add(str(v))
            codes[i, j] = c
    return codes, unseen
