"""HEVC partition geometry used by the layout tests (independent of oracle and product tables)."""


def decode_key(key):
    """TComDataCU.cpp:3388-3391: w + 100*(h + 100*(zIdx + 1000*(partSize + 10*depth + 100*partIdx)))."""
    w, key = key % 100, key // 100
    h, key = key % 100, key // 100
    z, key = key % 1000, key // 1000
    ps, key = key % 10, key // 10
    depth, part = key % 10, key // 10
    return w, h, z, ps, depth, part


def pu_rect(w, h, z, ps, depth, part):
    """HEVC geometry, independent of the oracle's table: CU origin from the z-order index (4x4 units,
    16x16 grid) and PU rectangle from PartSize (TypeDef.h enum: 2Nx2N,2NxN,Nx2N,NxN,2NxnU,2NxnD,nLx2N,nRx2N)."""
    assert w == h == 64 >> depth
    x4 = sum(((z >> (2 * b)) & 1) << b for b in range(4))
    y4 = sum(((z >> (2 * b + 1)) & 1) << b for b in range(4))
    cx, cy, S = 4 * x4, 4 * y4, w
    N, q = S // 2, S // 4
    table = {
        0: [(0, 0, S, S)],
        1: [(0, 0, S, N), (0, N, S, N)],
        2: [(0, 0, N, S), (N, 0, N, S)],
        4: [(0, 0, S, q), (0, q, S, S - q)],
        5: [(0, 0, S, S - q), (0, S - q, S, q)],
        6: [(0, 0, q, S), (q, 0, S - q, S)],
        7: [(0, 0, S - q, S), (S - q, 0, q, S)],
    }
    x, y, pw, ph = table[ps][part]
    return cx + x, cy + y, pw, ph
