"""Random alignment strings (DiffStr, diffstr.h:28-77) for the CIGAR tests."""
import numpy as np


def encode_columns(cols):
    """cols: string over '=XDI' (alignment columns) -> DiffStr bytes incl. the closing S byte and the 0.
    A run of matches longer than 62 is cut by M bytes (each stands for its count + 1 matches), the
    way alignment.c / diffstr.c write them."""
    out = bytearray()
    run = 0
    typ = {"X": 3, "D": 1, "I": 2}
    for c in cols:
        if c == "=":
            run += 1
            if run > 62:
                out.append((0 << 6) | 61)   # 61 matches + the M column itself = 62
                run -= 62
            continue
        out.append((typ[c] << 6) | run)
        run = 0
    out.append((3 << 6) | run)
    out.append(0)
    return bytes(out)


def random_columns(rng, n, p_x=0.03, p_d=0.01, p_i=0.01, gap_ext=0.4):
    cols = []
    while len(cols) < n:
        u = rng.random()
        if u < p_x:
            cols.append("X")
        elif u < p_x + p_d + p_i:
            g = "D" if u < p_x + p_d else "I"
            cols.append(g)
            while rng.random() < gap_ext:
                cols.append(g)
        else:
            cols.append("=")
    return "".join(cols)


def random_bytes_string(rng, n):
    """any non-zero bytes, closed by an S byte: strings the DP never writes (M bytes with small counts,
    adjacent gaps of different type, S bytes in a row) but the functions accept"""
    b = rng.integers(1, 256, n, dtype=np.uint8)
    b = bytes(int(x) for x in b) + bytes([(3 << 6) | int(rng.integers(0, 64))]) + b"\0"
    return b
