"""C++ host mirror (host/): builds, file formats and selection rules agree with numpy -- no GPU."""
import math
import os
import subprocess

import numpy as np
import pytest

from lia_ral_b200 import synth
from tests import lia_files as lf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "host", "build")


@pytest.fixture(scope="module")
def built():
    from lia_ral_b200 import capi
    capi.build()
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "host"), "-s", "-j", "8"])
    return BIN


def test_programs_exist_and_print_help(built):
    for p in ("TrainWorld", "TrainTarget", "ComputeTest", "IvExtractor", "TotalVariability", "IvTest", "IvNorm", "PLDA",
              "ComputeJFAStats", "EigenVoice", "EigenChannel", "EstimateDMatrix"):
        out = subprocess.run([os.path.join(built, p), "--help"], capture_output=True, text=True, timeout=60)
        assert out.returncode == 0 and p in out.stdout


def test_file_formats_and_selection(built, tmp_path):
    C, D0 = 8, 10
    w, mean, cov = synth.make_ubm(C, 9, seed=5)
    lf.write_raw_gmm(tmp_path / "wld.gmm", w, mean, cov)
    rng = np.random.default_rng(3)
    feats = {}
    for i, n in enumerate((57, 130)):
        X = rng.standard_normal((n, D0)).astype(np.float32)
        feats[f"f{i}"] = X
        (lf.write_spro4 if i == 0 else lf.write_spro4)(tmp_path / f"f{i}.prm", X)
    lf.write_lines(tmp_path / "f0.lbl", ["0.00 0.10 speech", "0.20 0.2999999 sil", "0.30 0.45 speech", "0.50 9.0 speech"])
    lf.write_lines(tmp_path / "f1.lbl", ["0.05 0.80 speech"])
    lf.write_lines(tmp_path / "ndx", [["f0", "spkA", "spkB"], ["f1", "spkA"]])
    lf.write_cfg(tmp_path / "t.cfg", mixtureFilesPath=str(tmp_path) + "/", loadMixtureFileExtension=".gmm",
                 loadMixtureFileFormat="RAW", featureFilesPath=str(tmp_path) + "/", loadFeatureFileExtension=".prm",
                 loadFeatureFileFormat="SPRO4", featureServerMask="0-3,5-9", labelFilesPath=str(tmp_path) + "/",
                 labelFilesExtension=".lbl", labelSelectedFrames="speech", frameLength=0.01,
                 inputWorldFilename="wld", ndxFilename=str(tmp_path / "ndx"), tmpPrefix=str(tmp_path / "tmp"))
    out = subprocess.run([os.path.join(built, "HostSelfTest"), "--config", str(tmp_path / "t.cfg")],
                         capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stdout + out.stderr
    r = {l.split()[0]: l.split()[1:] for l in out.stdout.strip().splitlines()}
    assert "FAILED" not in r
    g = r["gmm"]
    assert (int(g[0]), int(g[1])) == (C, 9)
    assert math.isclose(float(g[2]), w.sum(), rel_tol=1e-12) and math.isclose(float(g[3]), mean.sum(), rel_tol=1e-12)
    assert math.isclose(float(g[4]), cov.sum(), rel_tol=1e-12) and math.isclose(float(g[5]), 1.0, rel_tol=1e-12)
    assert float(r["xml_roundtrip"][0]) < 1e-14
    # normalizeMixture (N(0, 1) as a whole; meanOnly keeps the variances), reduceToTopWeights, computeMAPConst2
    nz = [float(v) for v in r["normalize"]]
    assert nz[0] < 1e-12 and nz[1] < 1e-12 and nz[2] == 0.0 and abs(nz[3]) < 1e-12
    keep = np.sort(np.argsort(-w, kind="stable")[:3])
    rd = r["reduce"]
    assert int(rd[0]) == 3 and math.isclose(float(rd[1]), 1.0, rel_tol=1e-12)
    assert np.allclose([float(v) for v in rd[2:5]], w[keep] / w[keep].sum(), rtol=1e-12)
    assert math.isclose(float(rd[5]), mean[keep[0], 0], rel_tol=1e-12)
    assert math.isclose(float(rd[6]), 1.0 / np.sqrt((2 * np.pi) ** 9 * np.prod(cov[keep[0]])), rel_tol=1e-10)
    assert float(r["map_const2"][0]) < 1e-12 and float(r["map_const2"][1]) == 0.0
    assert [int(v) for v in r["ndx"]] == [2, 5, 4]
    mask = [0, 1, 2, 3, 5, 6, 7, 8, 9]
    allx = np.concatenate([feats["f0"][:, mask], feats["f1"][:, mask]])
    assert (int(r["features"][0]), int(r["features"][1])) == (187, 9)
    assert math.isclose(float(r["features"][2]), float(allx.astype(np.float64).sum()), rel_tol=1e-9, abs_tol=1e-6)
    # label -> frames: end inclusive, clipped to the file, 0.2999999/0.01 rounds up to frame 30
    sel0 = list(range(0, 11)) + list(range(30, 46)) + list(range(50, 57))
    sel1 = list(range(5, 81))
    ref = feats["f0"][sel0][:, mask].astype(np.float64).sum() + feats["f1"][sel1][:, mask].astype(np.float64).sum()
    assert (int(r["selected"][0]), int(r["selected"][1])) == (4, len(sel0) + len(sel1))
    assert math.isclose(float(r["selected"][2]), float(ref), rel_tol=1e-9, abs_tol=1e-6)
    assert (int(r["matrix_roundtrip"][0]), int(r["matrix_roundtrip"][1])) == (3, 5) and float(r["matrix_roundtrip"][2]) < 1e-15
    assert int(r["bagged"][2]) <= 7 and 0 < int(r["bagged"][1]) < len(sel0) + len(sel1)
    assert math.isclose(float(r["setItParameter"][0]), 0.3) and int(r["setItParameter"][1]) == 30
    assert r["exception"] == ["1"]
    # IvTest score output (IvTest.cpp:412-465): segments outer, models inner, masked; binary variant
    lines = [l.split() for l in open(str(tmp_path / "tmp") + "_scores.res")]
    assert lines == [["F", "m0", "0", "s0", "-1"], ["F", "m1", "1", "s0", "0.5"], ["F", "m1", "1", "s1", "1"],
                     ["F", "m0", "0", "s2", "0"]]
    assert open(str(tmp_path / "tmp") + "_scores_model.txt").read().split() == ["m0", "m1"]
    assert open(str(tmp_path / "tmp") + "_scores_testSeg.txt").read().split() == ["s0", "s1", "s2"]
    assert r["scores_binary"] == ["2", "3", "1.5"]
    # rank sharding of NDX lines / segments (one process per GPU): contiguous, balanced, covering
    assert r["shard_range"] == ["0-4", "4-7", "7-10"]
    assert r["shard_weight"] == ["0-3", "3-4", "4-8"]      # cut after 1/3 and 2/3 of the 1800 frames
    # the same features through the other on-disk formats: HTK (big-endian by definition), SPRO3,
    # RAW and byte-swapped RAW (bigEndian) -- the FeatureServer block must be identical
    for fmt, ext, extra in (("HTK", ".htk", {}), ("SPRO3", ".sp3", {}), ("RAW", ".raw", {"vectSize": D0}),
                            ("RAW", ".rawbe", {"vectSize": D0, "bigEndian": "true"})):
        for name, X in feats.items():
            if fmt == "HTK":
                lf.write_htk(tmp_path / f"{name}{ext}", X)
            elif fmt == "SPRO3":
                lf.write_spro3(tmp_path / f"{name}{ext}", X)
            else:
                X.astype(">f4" if extra.get("bigEndian") else "<f4").tofile(tmp_path / f"{name}{ext}")
        out2 = subprocess.run([os.path.join(built, "HostSelfTest"), "--config", str(tmp_path / "t.cfg"),
                               "--loadFeatureFileFormat", fmt, "--loadFeatureFileExtension", ext] +
                              [a for k, v in extra.items() for a in (f"--{k}", str(v))],
                              capture_output=True, text=True, timeout=60)
        r2 = {l.split()[0]: l.split()[1:] for l in out2.stdout.strip().splitlines()}
        assert r2["features"] == r["features"] and r2["selected"] == r["selected"], (fmt, ext, out2.stdout)


def test_unimplemented_modes_are_refused_loudly(built, tmp_path):
    """Options of the reference this engine does not implement must be refused with a message, never silently ignored
    (both refusals happen before any engine call, so this runs without a GPU)."""
    w, mean, cov = synth.make_ubm(4, 5, seed=7)
    lf.write_raw_gmm(tmp_path / "wld.gmm", w, mean, cov)
    lf.write_lines(tmp_path / "ids", [["spk", "f0"]])
    lf.write_lines(tmp_path / "ndx", [["f0", "spk"]])
    common = dict(mixtureFilesPath=str(tmp_path) + "/", loadMixtureFileExtension=".gmm", loadMixtureFileFormat="RAW",
                  inputWorldFilename="wld", labelSelectedFrames="speech", targetIdList=str(tmp_path / "ids"),
                  ndxFilename=str(tmp_path / "ndx"), outputFilename=str(tmp_path / "out.res"), gender="F",
                  MAPAlgo="MAPOccDep", meanAdapt="true", MAPRegFactorMean=14.0)
    lf.write_cfg(tmp_path / "r.cfg", **common)
    out = subprocess.run([os.path.join(built, "TrainTarget"), "--config", str(tmp_path / "r.cfg"), "--useModelData", "true"],
                         capture_output=True, text=True, timeout=60)
    assert "not implemented" in out.stdout and not os.path.exists(tmp_path / "spk.gmm")
    out = subprocess.run([os.path.join(built, "TrainTarget"), "--config", str(tmp_path / "r.cfg"), "--MAPAlgo", "MLLR"],
                         capture_output=True, text=True, timeout=60)
    assert "not implemented" in out.stdout
    out = subprocess.run([os.path.join(built, "ComputeTest"), "--config", str(tmp_path / "r.cfg"),
                          "--channelCompensation", "NAP"], capture_output=True, text=True, timeout=60)
    assert out.returncode == 1 and "not implemented" in out.stdout
    out = subprocess.run([os.path.join(built, "EigenChannel"), "--config", str(tmp_path / "r.cfg"),
                          "--eigenChannelMode", "XFA"], capture_output=True, text=True, timeout=60)
    assert out.returncode == 1 and "wrong eigenChannelMode" in out.stdout
