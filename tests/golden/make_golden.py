#!/usr/bin/env python
"""Derives the committed golden vectors from the reference's own intact test fixtures.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py
Outputs (small, committed):
    gmmtokenizer.npz   LIA_Utils/GmmTokenizer/test/{wld,test1.prm,test1.lbl,test1.sym.ref,
                       mce_matrix.mat.ref}: RAW GMM (128x32), the 37 label-selected frames after
                       featureServerMask 0-15,17-32, the expected best-Gaussian stream and the
                       top-20 confusion matrix (the .ref was produced with topDistribsCount 20).
    wld_validate.npz   LIA_SpkDet/TrainWorld/test/wld.validate (XML GMM 10x32): weight/cst/det/
                       covInv/mean as printed -> computeAll identity KAT.
    traintarget.npz    LIA_SpkDet/TrainTarget/test/{wld,test1.prm,test1.lbl,test1.validate.gmm}: RAW world GMM
                       (128x32), the 32 label-selected frames, and the means of the reference's adapted
                       model.  The .validate.gmm is DAMAGED (68743 bytes, one byte of component 65's record is
                       missing): components 0..64 are read in place, 66..127 one byte earlier (their covInv
                       blocks then equal the world's bit for bit, which is how the cut was located), component
                       65 is dropped.  An INDICATIVE pin of the posterior / EM / MAPOccDep semantics: one
                       MAP iteration (MAPRegFactorMean 10) from the world model reproduces these means to
                       1e-3 absolute (the fixture's values sit on a 1.8e-4 grid).
Only DATA is converted; no reference source is copied.
"""
import math
import os
import re
import struct

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def read_raw_gmm(path):
    raw = open(path, "rb").read()
    C, D = struct.unpack("<II", raw[:8])
    w = np.frombuffer(raw, "<f8", C, 8).copy()
    off = 8 + 8 * C
    cst, det = np.zeros(C), np.zeros(C)
    ci, mu = np.zeros((C, D)), np.zeros((C, D))
    for c in range(C):
        cst[c], det[c] = struct.unpack("<dd", raw[off:off + 16])
        off += 17  # + 1 flag byte
        ci[c] = np.frombuffer(raw, "<f8", D, off)
        off += 8 * D
        mu[c] = np.frombuffer(raw, "<f8", D, off)
        off += 8 * D
    assert off == len(raw)
    return w, cst, det, ci, mu


def parse_mask(s):
    out = []
    for part in s.split(","):
        if "-" in part:
            a, b = part.split("-")
            out += list(range(int(a), int(b) + 1))
        else:
            out.append(int(part))
    return out


def time_to_frame(t, frame_length):
    # SegTools.cpp:135-142 timeToFrameIdx: floor, except a fractional part > 0.99999 rounds up
    frac, whole = math.modf(t / frame_length)
    return int(whole) + 1 if frac > 0.99999 else int(whole)


def gmmtokenizer():
    d = os.path.join(REF, "LIA_Utils/GmmTokenizer/test")
    w, cst, det, ci, mu = read_raw_gmm(os.path.join(d, "wld"))
    prm = open(os.path.join(d, "test1.prm"), "rb").read()
    hdr = struct.unpack("<4I", prm[:16])
    x = np.frombuffer(prm, "<f4", -1, 16).reshape(hdr[2], -1)
    x = x[:, parse_mask("0-15,17-32")].copy()
    sel = []
    for line in open(os.path.join(d, "test1.lbl")):
        b, e, lab = line.split()
        if lab != "male":
            continue
        # segment end is inclusive (SegTools.cpp:269-270)
        sel += list(range(time_to_frame(float(b), 0.01), time_to_frame(float(e), 0.01) + 1))
    sym = np.array(open(os.path.join(d, "test1.sym.ref")).read().split(), dtype=np.int64)
    mce = np.loadtxt(os.path.join(d, "mce_matrix.mat.ref"), skiprows=1).astype(np.int64)
    np.savez_compressed(os.path.join(HERE, "gmmtokenizer.npz"), w=w, cst=cst, det=det, covinv=ci,
                        mean=mu, frames=x, selected=np.array(sel, dtype=np.int64), sym_ref=sym,
                        mce_ref=mce, mce_topk=np.int64(20))
    print("gmmtokenizer:", x.shape, len(sel), sym, mce.sum())


def wld_validate():
    txt = open(os.path.join(REF, "LIA_SpkDet/TrainWorld/test/wld.validate")).read()
    head = re.search(r'distribCount="(\d+)" vectSize="(\d+)"', txt)
    C, D = int(head.group(1)), int(head.group(2))
    w, cst, det = np.zeros(C), np.zeros(C), np.zeros(C)
    ci, mu = np.zeros((C, D)), np.zeros((C, D))
    blocks = re.findall(r'<DistribGD i="(\d+)" weight="([^"]+)" cst="([^"]+)" det="([^"]+)">(.*?)'
                        r'</DistribGD>', txt, flags=re.S)
    assert len(blocks) == C
    for i, ws, cs, ds, body in blocks:
        i = int(i)
        w[i], cst[i], det[i] = float(ws), float(cs), float(ds)
        for j, v in re.findall(r'<covInv i="(\d+)">([^<]+)</covInv>', body):
            ci[i, int(j)] = float(v)
        for j, v in re.findall(r'<mean i="(\d+)">([^<]+)</mean>', body):
            mu[i, int(j)] = float(v)
    np.savez_compressed(os.path.join(HERE, "wld_validate.npz"), w=w, cst=cst, det=det, covinv=ci,
                        mean=mu)
    print("wld_validate:", C, D, w.sum())


def traintarget():
    d = os.path.join(REF, "LIA_SpkDet/TrainTarget/test")
    w, cst, det, ci, mu = read_raw_gmm(os.path.join(d, "wld"))
    raw = open(os.path.join(d, "test1.validate.gmm"), "rb").read()
    C, D = struct.unpack("<II", raw[:8])
    rec, off0 = 17 + 16 * D, 8 + 8 * C
    assert len(raw) == off0 + C * rec - 1          # the damaged fixture: one byte short
    mean_ref = np.full((C, D), np.nan)
    ok = np.ones(C, dtype=bool)
    for c in range(C):
        if c == 65:
            ok[c] = False
            continue
        o = off0 + c * rec - (0 if c < 65 else 1)
        assert np.array_equal(np.frombuffer(raw, "<f8", D, o + 17), ci[c])   # only the means are adapted
        mean_ref[c] = np.frombuffer(raw, "<f8", D, o + 17 + 8 * D)
    prm = open(os.path.join(d, "test1.prm"), "rb").read()
    hdr = struct.unpack("<4I", prm[:16])
    x = np.frombuffer(prm, "<f4", -1, 16).reshape(hdr[2], -1)[:, parse_mask("0-15,17-32")].copy()
    sel = []
    for line in open(os.path.join(d, "test1.lbl")):
        b, e, lab = line.split()
        if lab == "speech":
            sel += list(range(time_to_frame(float(b), 0.01), time_to_frame(float(e), 0.01) + 1))
    np.savez_compressed(os.path.join(HERE, "traintarget.npz"), w=w, cst=cst, covinv=ci, mean=mu, frames=x,
                        selected=np.array(sel, dtype=np.int64), mean_ref=mean_ref, ok=ok,
                        map_reg_factor_mean=np.float64(10.0))
    print("traintarget:", x.shape, len(sel), int(ok.sum()), np.nanmax(np.abs(mean_ref - mu)))


if __name__ == "__main__":
    gmmtokenizer()
    wld_validate()
    traintarget()
