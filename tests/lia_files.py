"""Writers for the ALIZE / LIA_RAL on-disk formats used by the host-layer tests."""
import struct

import numpy as np


def write_raw_gmm(path, w, mean, cov):
    """uint32 C, uint32 D, double w[C], then per component cst, det, 1 flag byte, covInv[D], mean[D]"""
    C, D = mean.shape
    with open(path, "wb") as f:
        f.write(struct.pack("<II", C, D))
        f.write(np.asarray(w, "<f8").tobytes())
        for c in range(C):
            det = float(np.prod(cov[c]))
            cst = 1.0 / ((2 * np.pi) ** (D / 2) * np.sqrt(det))
            f.write(struct.pack("<dd", cst, det))
            f.write(b"\x00")
            f.write((1.0 / cov[c]).astype("<f8").tobytes())
            f.write(mean[c].astype("<f8").tobytes())


def read_raw_gmm(path):
    raw = open(path, "rb").read()
    C, D = struct.unpack("<II", raw[:8])
    w = np.frombuffer(raw, "<f8", C, 8).copy()
    off = 8 + 8 * C
    ci, mu = np.zeros((C, D)), np.zeros((C, D))
    for c in range(C):
        off += 17
        ci[c] = np.frombuffer(raw, "<f8", D, off)
        off += 8 * D
        mu[c] = np.frombuffer(raw, "<f8", D, off)
        off += 8 * D
    return w, mu, 1.0 / ci


def write_spro4(path, X):
    """uint16 dim, uint32 flags, float rate, then frames x dim float32 (no text header)"""
    with open(path, "wb") as f:
        f.write(struct.pack("<HIf", X.shape[1], 0, 100.0))
        f.write(np.ascontiguousarray(X, "<f4").tobytes())


def write_spro3(path, X):
    with open(path, "wb") as f:
        f.write(struct.pack("<4I", 2, 16, X.shape[0], 9))
        f.write(np.ascontiguousarray(X, "<f4").tobytes())


def write_htk(path, X, period=100000, kind=9):
    """HTK parameter file: 12-byte big-endian header + big-endian float32 (kind 9 = USER)."""
    X = np.ascontiguousarray(X, dtype=np.float32)
    with open(path, "wb") as f:
        f.write(struct.pack(">iihh", X.shape[0], period, 4 * X.shape[1], kind))
        f.write(X.astype(">f4").tobytes())


def write_db(path, M):
    M = np.atleast_2d(np.asarray(M, dtype="<f8"))
    with open(path, "wb") as f:
        f.write(struct.pack("<II", *M.shape))
        f.write(np.ascontiguousarray(M).tobytes())


def read_db(path):
    raw = open(path, "rb").read()
    r, c = struct.unpack("<II", raw[:8])
    return np.frombuffer(raw, "<f8", r * c, 8).reshape(r, c).copy()


def write_cfg(path, **kv):
    with open(path, "w") as f:
        f.write("*** generated test configuration ***\n")
        for k, v in kv.items():
            f.write(f"{k}\t{v}\n")


def write_lines(path, lines):
    with open(path, "w") as f:
        for l in lines:
            f.write((" ".join(map(str, l)) if not isinstance(l, str) else l) + "\n")
