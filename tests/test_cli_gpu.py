"""The five command-line programs of the C++ host mirror (host/build/*), end to end on files in the
reference's formats, against the oracle: same configs keys, same outputs (NIST result lines, RAW
GMM, DB matrices, per-id i-vector files)."""
import os
import subprocess
import time

import numpy as np
import pytest

from lia_ral_b200 import synth
from tests import lia_files as lf

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "host", "build")


def _run(prog, cfg, **over):
    cmd = [os.path.join(BIN, prog), "--config", str(cfg)]
    for k, v in over.items():
        cmd += [f"--{k}", str(v)]
    t0 = time.perf_counter()
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    if os.environ.get("LIA_CLI_TIMES"):   # per-program wall times for profiling the suite itself
        with open(os.environ["LIA_CLI_TIMES"], "a") as f:
            f.write(f"{prog} {time.perf_counter() - t0:.2f}s\n")
    assert out.returncode == 0, out.stdout + out.stderr
    assert "Exception" not in out.stdout, out.stdout
    return out.stdout


@pytest.fixture(scope="module")
def world(tmp_path_factory, oracle):
    if not os.path.exists(os.path.join(BIN, "TrainWorld")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "host"), "-s", "-j", "8"])
    d = tmp_path_factory.mktemp("cli")
    C, D = 32, 12
    w, mean, cov = synth.make_ubm(C, D, seed=51)
    lf.write_raw_gmm(d / "wld.gmm", w, mean, cov)
    utts = {}
    for i in range(6):
        X = synth.make_frames(w, mean, cov * 2.0, 300 + 40 * i, seed=60 + i)
        utts[f"utt{i}"] = X
        lf.write_spro4(d / f"utt{i}.prm", X)
        # label files: two speech segments with a gap; utt5 has no label file (default label)
        if i < 5:
            lf.write_lines(d / f"utt{i}.lbl", [f"0.10 {1.5 + 0.1 * i:.2f} speech", "1.80 1.95 noise", f"2.00 {2.6 + 0.1 * i:.2f} speech"])
    common = dict(mixtureFilesPath=str(d) + "/", loadMixtureFileExtension=".gmm", saveMixtureFileExtension=".gmm",
                  loadMixtureFileFormat="RAW", saveMixtureFileFormat="RAW", featureFilesPath=str(d) + "/",
                  loadFeatureFileExtension=".prm", loadFeatureFileFormat="SPRO4", labelFilesPath=str(d) + "/",
                  labelFilesExtension=".lbl", labelSelectedFrames="speech", addDefaultLabel="true",
                  defaultLabel="speech", frameLength=0.01, matrixFilesPath=str(d) + "/", loadMatrixFormat="DB",
                  saveMatrixFormat="DB", loadMatrixFilesExtension=".mat", saveMatrixFilesExtension=".mat",
                  minLLK=-200, maxLLK=200)
    return dict(dir=d, C=C, D=D, w=w, mean=mean, cov=cov, utts=utts, common=common)


def _selected(name, X):
    """frames selected by the label rule (end inclusive, clipped)"""
    i = int(name[3:])
    if i == 5:
        return np.arange(len(X))
    a = np.arange(10, int(round((1.5 + 0.1 * i) * 100)) + 1)
    b = np.arange(200, int(round((2.6 + 0.1 * i) * 100)) + 1)
    return np.concatenate([a, b[b < len(X)]])


def test_compute_test_cli(world, oracle):
    d = world["dir"]
    clients = {}
    for k in range(3):
        p = synth.perturb_ubm(world["w"], world["mean"], world["cov"], seed=70 + k, frac=0.4, scale=0.5)
        clients[f"spk{k}"] = p
        lf.write_raw_gmm(d / f"spk{k}.gmm", *p)
    lf.write_lines(d / "test.ndx", [["utt0", "spk0", "spk1"], ["utt3", "spk2"], ["utt5", "spk0", "spk1", "spk2"]])
    lf.write_cfg(d / "ct.cfg", **world["common"], ndxFilename=str(d / "test.ndx"), inputWorldFilename="wld",
                 outputFilename=str(d / "ct.res"), gender="F", topDistribsCount=5,
                 computeLLKWithTopDistribs="COMPLETE")
    _run("ComputeTest", d / "ct.cfg")
    lines = [l.split() for l in open(d / "ct.res")]
    assert [(l[1], l[3]) for l in lines] == [("spk0", "utt0"), ("spk1", "utt0"), ("spk2", "utt3"),
                                             ("spk0", "utt5"), ("spk1", "utt5"), ("spk2", "utt5")]
    ow = oracle.gmm(world["w"], world["mean"], world["cov"])
    for l in lines:
        X = np.ascontiguousarray(world["utts"][l[3]][_selected(l[3], world["utts"][l[3]])])
        llk_w, idx, _, rest, _ = oracle.llk_determine_top(ow, X, 5, True)
        llk_c = oracle.llk_use_top(oracle.gmm(*clients[l[1]]), X, idx, rest, True)
        ref = llk_c.mean() - llk_w.mean()
        assert l[0] == "F" and int(l[2]) == int(ref > 0)
        assert abs(float(l[4]) - ref) < 2e-4
    # segmental mode: one line per selected segment with begin / end times
    _run("ComputeTest", d / "ct.cfg", segmentLLR="true", outputFilename=str(d / "ct_seg.res"))
    seg_lines = [l.split() for l in open(d / "ct_seg.res")]
    assert len(seg_lines) == 2 * 2 + 2 * 1 + 1 * 3 and all(len(l) == 7 for l in seg_lines)


def test_train_world_cli(world, oracle):
    d = world["dir"]
    start = synth.perturb_ubm(world["w"], world["mean"], world["cov"], seed=81, frac=1.0, scale=0.4)
    lf.write_raw_gmm(d / "start.gmm", *start)
    lf.write_lines(d / "train.lst", [[f"utt{i}"] for i in range(6)])
    lf.write_cfg(d / "tw.cfg", **world["common"], inputFeatureFilename=str(d / "train.lst"),
                 inputWorldFilename="start", outputWorldFilename="trained", nbTrainIt=3,
                 baggedFrameProbability=1.0, initVarianceFlooring=0.4, finalVarianceFlooring=0.2,
                 initVarianceCeiling=8.0, finalVarianceCeiling=6.0)
    _run("TrainWorld", d / "tw.cfg")
    w, mean, cov = lf.read_raw_gmm(d / "trained.gmm")
    X = np.concatenate([world["utts"][f"utt{i}"][_selected(f"utt{i}", world["utts"][f"utt{i}"])] for i in range(6)])
    X = np.ascontiguousarray(X)
    _, gcov = oracle.mean_cov(X)
    g = oracle.gmm(*start)
    for it in range(3):
        fl = oracle.set_it_parameter(0.4, 0.2, 3, it)
        ce = oracle.set_it_parameter(8.0, 6.0, 3, it)
        _, _, occ, m1, m2 = oracle.em_accumulate(g, X)
        wn, mn, cn = oracle.em_get(g, occ, m1, m2)
        cn, _, _ = oracle.variance_control(cn, fl, ce, gcov)
        g = oracle.gmm(wn, mn, cn)
    assert np.allclose(w, g.w, rtol=1e-3, atol=1e-6)
    assert np.abs(mean - g.mean).max() < 1e-3 * np.abs(g.mean).max()
    assert np.abs(cov - g.cov).max() < 2e-3 * np.abs(g.cov).max()


def test_train_world_reduction_and_normalize_cli(world, oracle):
    """componentReduction / targetMixtureDistribCount and normalizeModel of trainModelStream (TrainTools.cpp:1078-1098):
    after getEM + varianceControl the heaviest nbTop components survive (nbTop shrinks linearly to the target, weights
    renormalised), then the mixture is re-expressed relative to its moment-matched single Gaussian
    (normalizeMixture :287-315 with the N(0, 1) target)."""
    d = world["dir"]
    start = synth.perturb_ubm(world["w"], world["mean"], world["cov"], seed=81, frac=1.0, scale=0.4)
    lf.write_raw_gmm(d / "start_r.gmm", *start)
    lf.write_lines(d / "train_r.lst", [[f"utt{i}"] for i in range(6)])
    X = np.ascontiguousarray(np.concatenate(
        [world["utts"][f"utt{i}"][_selected(f"utt{i}", world["utts"][f"utt{i}"])] for i in range(6)]))
    _, gcov = oracle.mean_cov(X)
    C0, target, nb_it = world["C"], 20, 3

    def reference(reduce, normalize, mean_only=False, norm_it=1):
        g = oracle.gmm(*start)
        for it in range(nb_it):
            fl = oracle.set_it_parameter(0.4, 0.2, nb_it, it)
            ce = oracle.set_it_parameter(8.0, 6.0, nb_it, it)
            _, _, occ, m1, m2 = oracle.em_accumulate(g, X)
            wn, mn, cn = oracle.em_get(g, occ, m1, m2)
            cn, _, _ = oracle.variance_control(cn, fl, ce, gcov)
            if reduce:
                nb_top = C0 - int((it + 1) * ((C0 - target) / nb_it))
                if it == nb_it - 1:
                    nb_top = target
                if nb_top < len(wn):
                    keep = np.sort(np.argsort(-wn, kind="stable")[:nb_top])
                    wn, mn, cn = wn[keep] / wn[keep].sum(), mn[keep], cn[keep]
            if normalize:
                for _ in range(norm_it):
                    gm = (wn[:, None] * mn).sum(0) / wn.sum()
                    gc = (wn[:, None] * (cn + mn ** 2)).sum(0) / wn.sum() - gm ** 2
                    mn = (mn - gm) / np.sqrt(gc)
                    if not mean_only:
                        cn = cn / gc
            g = oracle.gmm(wn, mn, cn)
        return g

    base = dict(**world["common"], inputFeatureFilename=str(d / "train_r.lst"), inputWorldFilename="start_r",
                nbTrainIt=nb_it, baggedFrameProbability=1.0, initVarianceFlooring=0.4, finalVarianceFlooring=0.2,
                initVarianceCeiling=8.0, finalVarianceCeiling=6.0)
    cases = [("red", dict(componentReduction="true", targetMixtureDistribCount=target), (True, False)),
             ("norm", dict(normalizeModel="true"), (False, True)),
             ("both", dict(componentReduction="true", targetMixtureDistribCount=target, normalizeModel="true",
                           normalizeModelMeanOnly="true", normalizeModelNbIt=2), (True, True, True, 2))]
    for name, extra, ref_args in cases:
        lf.write_cfg(d / f"tw_{name}.cfg", **base, outputWorldFilename=f"trained_{name}", **extra)
        _run("TrainWorld", d / f"tw_{name}.cfg")
        w, mean, cov = lf.read_raw_gmm(d / f"trained_{name}.gmm")
        g = reference(*ref_args)
        assert len(w) == len(g.w) == (target if ref_args[0] else C0)
        assert abs(w.sum() - 1.0) < 1e-9
        assert np.allclose(w, g.w, rtol=1e-3, atol=1e-6)
        assert np.abs(mean - g.mean).max() < 1e-3 * np.abs(g.mean).max()
        assert np.abs(cov - g.cov).max() < 2e-3 * np.abs(g.cov).max()
        if ref_args[1] and not (len(ref_args) > 2 and ref_args[2]):
            # the normalised mixture has zero mean and unit variance as a whole
            assert np.abs((w[:, None] * mean).sum(0)).max() < 1e-9
            assert np.abs((w[:, None] * (cov + mean ** 2)).sum(0) - 1.0).max() < 1e-9


def test_train_world_multi_stream_cli(world, oracle):
    """inputStreamList / weightStreamList (TrainWorld.cpp:120-141, trainModelStream TrainTools.cpp:1030-1110).
    Weights (1, 0) make the bagging deterministic: stream A is taken whole (probability exactly 1), stream B
    never -- while the global covariance of the variance control still covers both streams."""
    d = world["dir"]
    start = synth.perturb_ubm(world["w"], world["mean"], world["cov"], seed=82, frac=1.0, scale=0.4)
    lf.write_raw_gmm(d / "start2.gmm", *start)
    lf.write_lines(d / "stream_a.lst", [[f"utt{i}"] for i in range(3)])
    lf.write_lines(d / "stream_b.lst", [[f"utt{i}"] for i in range(3, 6)])
    lf.write_lines(d / "streams.lst", [[str(d / "stream_a.lst")], [str(d / "stream_b.lst")]])
    lf.write_lines(d / "weights.lst", [["1.0"], ["0.0"]])
    lf.write_cfg(d / "tw2.cfg", **world["common"], inputStreamList=str(d / "streams.lst"),
                 weightStreamList=str(d / "weights.lst"), inputWorldFilename="start2", outputWorldFilename="trained2",
                 nbTrainIt=2, baggedFrameProbability=1.0, initVarianceFlooring=0.4, finalVarianceFlooring=0.2,
                 initVarianceCeiling=8.0, finalVarianceCeiling=6.0)
    _run("TrainWorld", d / "tw2.cfg")
    w, mean, cov = lf.read_raw_gmm(d / "trained2.gmm")
    sel = lambda i: world["utts"][f"utt{i}"][_selected(f"utt{i}", world["utts"][f"utt{i}"])]
    Xa = np.ascontiguousarray(np.concatenate([sel(i) for i in range(3)]))
    Xall = np.ascontiguousarray(np.concatenate([sel(i) for i in range(6)]))
    _, gcov = oracle.mean_cov(Xall)
    g = oracle.gmm(*start)
    for it in range(2):
        fl = oracle.set_it_parameter(0.4, 0.2, 2, it)
        ce = oracle.set_it_parameter(8.0, 6.0, 2, it)
        _, _, occ, m1, m2 = oracle.em_accumulate(g, Xa)
        wn, mn, cn = oracle.em_get(g, occ, m1, m2)
        cn, _, _ = oracle.variance_control(cn, fl, ce, gcov)
        g = oracle.gmm(wn, mn, cn)
    assert np.allclose(w, g.w, rtol=1e-3, atol=1e-6)
    assert np.abs(mean - g.mean).max() < 1e-3 * np.abs(g.mean).max()
    assert np.abs(cov - g.cov).max() < 2e-3 * np.abs(g.cov).max()
    # equal weights (the default): both streams are bagged at random -- the run completes and moves the model
    lf.write_cfg(d / "tw3.cfg", **world["common"], inputStreamList=str(d / "streams.lst"),
                 inputWorldFilename="start2", outputWorldFilename="trained3", nbTrainIt=2,
                 baggedFrameProbability=0.8, initVarianceFlooring=0.4, finalVarianceFlooring=0.2,
                 initVarianceCeiling=8.0, finalVarianceCeiling=6.0)
    _run("TrainWorld", d / "tw3.cfg")
    w3, mean3, _ = lf.read_raw_gmm(d / "trained3.gmm")
    assert abs(w3.sum() - 1.0) < 1e-9 and np.abs(mean3 - start[1]).max() > 1e-3


def test_compute_jfa_stats_cli(world, oracle):
    """ComputeJFAStats (ComputeJFAStats.cpp:71-87): N, F_X per speaker (NDX line) and N_h, F_X_h per session
    (every element of a line is a session file, JFATranslate)."""
    d, C, D = world["dir"], world["C"], world["D"]
    ndx = [["utt0", "utt1"], ["utt2"], ["utt3", "utt4", "utt5"]]
    lf.write_lines(d / "jfa.ndx", ndx)
    lf.write_cfg(d / "jfa.cfg", **world["common"], ndxFilename=str(d / "jfa.ndx"), inputWorldFilename="wld",
                 nullOrderStatSpeaker="N_jfa", firstOrderStatSpeaker="FX_jfa", nullOrderStatSession="Nh_jfa",
                 firstOrderStatSession="FXh_jfa")
    _run("ComputeJFAStats", d / "jfa.cfg")
    N, F = lf.read_db(d / "N_jfa.mat"), lf.read_db(d / "FX_jfa.mat")
    Nh, Fh = lf.read_db(d / "Nh_jfa.mat"), lf.read_db(d / "FXh_jfa.mat")
    assert N.shape == (3, C) and F.shape == (3, C * D) and Nh.shape == (6, C) and Fh.shape == (6, C * D)
    ow = oracle.gmm(world["w"], world["mean"], world["cov"])
    session = 0
    for spk_i, line in enumerate(ndx):
        n_spk, f_spk = np.zeros(C), np.zeros(C * D)
        for u in line:
            X = np.ascontiguousarray(world["utts"][u][_selected(u, world["utts"][u])])
            n1, f1 = oracle.bwstats(ow, X, np.zeros(len(X), dtype=np.int32), 1)
            assert np.abs(Nh[session] - n1[0]).max() < 1e-4 * np.abs(n1).max()
            assert np.abs(Fh[session] - f1[0]).max() < 1e-4 * np.abs(f1).max()
            n_spk += n1[0]
            f_spk += f1[0]
            session += 1
        assert np.abs(N[spk_i] - n_spk).max() < 1e-4 * np.abs(n_spk).max()
        assert np.abs(F[spk_i] - f_spk).max() < 1e-4 * np.abs(f_spk).max()


def test_ivector_and_tv_cli(world, oracle):
    d, C, D, R = world["dir"], world["C"], world["D"], 5
    invvar = (1.0 / world["cov"]).reshape(-1)
    T = synth.make_T(R, C, D, invvar, seed=91, scale=0.05)
    lf.write_db(d / "TV.mat", T)
    ids = [["idA", "utt0", "utt1"], ["idB", "utt2"], ["idC", "utt3", "utt4", "utt0"]]  # utt0 on two lines
    lf.write_lines(d / "ids.ndx", ids)
    os.makedirs(d / "iv", exist_ok=True)
    lf.write_cfg(d / "iv.cfg", **world["common"], targetIdList=str(d / "ids.ndx"), inputWorldFilename="wld",
                 totalVariabilityNumber=R, totalVariabilityMatrix="TV", nullOrderStatSpeaker="N_iv",
                 firstOrderStatSpeaker="F_iv", saveVectorFilesPath=str(d / "iv") + "/", vectorFilesExtension=".y")
    _run("IvExtractor", d / "iv.cfg")
    ow = oracle.gmm(world["w"], world["mean"], world["cov"])
    N = np.zeros((3, C))
    F = np.zeros((3, C * D))
    for row, line in enumerate(ids):
        for u in line[1:]:
            X = np.ascontiguousarray(world["utts"][u][_selected(u, world["utts"][u])])
            n1, f1 = oracle.bwstats(ow, X, np.zeros(len(X), dtype=np.int32), 1)
            N[row] += n1[0]
            F[row] += f1[0]
    assert np.allclose(lf.read_db(d / "N_iv.mat"), N, rtol=1e-4, atol=1e-6)
    Fc = oracle.tv_subtract_m(N, F, world["mean"].reshape(-1))
    tett = oracle.tv_tett(T, invvar, C, D)
    W = oracle.tv_ivectors(N, Fc, T, invvar, tett)
    for row, line in enumerate(ids):
        y = lf.read_db(d / "iv" / f"{line[0]}.y")
        assert y.shape == (1, R) and np.abs(y[0] - W[row]).max() < 1e-4 * np.abs(W).max()
    # TotalVariability: same lists without ids, statistics recomputed, 2 EM iterations + minDivergence
    lf.write_lines(d / "tv.ndx", [l[1:] for l in ids])
    lf.write_cfg(d / "tv.cfg", **world["common"], ndxFilename=str(d / "tv.ndx"), inputWorldFilename="wld",
                 totalVariabilityNumber=R, totalVariabilityMatrix="TV_out", loadInitTotalVariabilityMatrix="true",
                 initTotalVariabilityMatrix="TV", nullOrderStatSpeaker="N_tv", firstOrderStatSpeaker="F_tv",
                 nbIt=2, minDivergence="true", meanEstimate="meanEst")
    _run("TotalVariability", d / "tv.cfg")
    Tr, mean = T.copy(), world["mean"].reshape(-1).copy()
    n_sessions = sum(len(l) - 1 for l in ids)
    for it in range(2):
        Fc = oracle.tv_subtract_m(N, F, mean)
        tett = oracle.tv_tett(Tr, invvar, C, D)
        _, A, Cmx, Rm, r, mw = oracle.tv_estep(N, Fc, Tr, invvar, tett)
        Tr = oracle.tv_mstep(A, Cmx, C, D)
        mean, Tr = oracle.tv_mindiv(Rm, r, mw, mean, Tr, float(n_sessions), C, D)
    got = lf.read_db(d / "TV_out.mat")
    assert np.abs(got - Tr).max() < 1e-3 * np.abs(Tr).max()
    assert np.abs(lf.read_db(d / "meanEst.mat")[0] - mean).max() < 1e-4 * np.abs(mean).max()


def test_eigenvoice_cli(world, oracle):
    """EigenVoice (EigenVoice.cpp:71-165): with D = Z = U = X = 0 the JFAAcc iteration IS the TVAcc iteration on the
    per-speaker statistics (estimateVEVT / estimateAndInverseL_EV / estimateYandV / updateVestimate), so V after two
    iterations is the oracle's T-matrix EM without minDivergence; the program also leaves the four JFA statistics."""
    d, C, D, R = world["dir"], world["C"], world["D"], 5
    invvar = (1.0 / world["cov"]).reshape(-1)
    V0 = synth.make_T(R, C, D, invvar, seed=291, scale=0.05)
    lf.write_db(d / "V0.mat", V0)
    ndx = [["utt0", "utt1"], ["utt2"], ["utt3", "utt4"], ["utt5"]]
    lf.write_lines(d / "ev.ndx", ndx)
    lf.write_cfg(d / "ev.cfg", **world["common"], ndxFilename=str(d / "ev.ndx"), inputWorldFilename="wld",
                 eigenVoiceNumber=R, eigenVoiceMatrix="V_out", loadInitEigenVoiceMatrix="true", initEigenVoiceMatrix="V0",
                 nullOrderStatSpeaker="N_ev", firstOrderStatSpeaker="FX_ev", nullOrderStatSession="Nh_ev",
                 firstOrderStatSession="FXh_ev", nbIt=2)
    _run("EigenVoice", d / "ev.cfg")
    ow = oracle.gmm(world["w"], world["mean"], world["cov"])
    N, F = np.zeros((len(ndx), C)), np.zeros((len(ndx), C * D))
    for row, line in enumerate(ndx):
        for u in line:
            X = np.ascontiguousarray(world["utts"][u][_selected(u, world["utts"][u])])
            n1, f1 = oracle.bwstats(ow, X, np.zeros(len(X), dtype=np.int32), 1)
            N[row] += n1[0]
            F[row] += f1[0]
    assert lf.read_db(d / "Nh_ev.mat").shape == (6, C) and lf.read_db(d / "FXh_ev.mat").shape == (6, C * D)
    assert np.abs(lf.read_db(d / "N_ev.mat") - N).max() < 1e-4 * np.abs(N).max()
    Vr, mean = V0.copy(), world["mean"].reshape(-1)
    for it in range(2):
        Fc = oracle.tv_subtract_m(N, F, mean)
        tett = oracle.tv_tett(Vr, invvar, C, D)
        _, A, Cmx, _, _, _ = oracle.tv_estep(N, Fc, Vr, invvar, tett)
        Vr = oracle.tv_mstep(A, Cmx, C, D)
    got = lf.read_db(d / "V_out.mat")
    assert np.abs(got - Vr).max() < 1e-3 * np.abs(Vr).max()


def test_eigenchannel_cli(world, oracle):
    """EigenChannel, JFA mode (EigenChannel.cpp:71-160): speaker factors with V, session statistics minus
    N_h o (M + V y_spk), then the channel matrix U by the TVAcc iteration on the session statistics."""
    d, C, D, Rv, Ru = world["dir"], world["C"], world["D"], 4, 3
    invvar = (1.0 / world["cov"]).reshape(-1)
    V0 = synth.make_T(Rv, C, D, invvar, seed=391, scale=0.05)
    U0 = synth.make_T(Ru, C, D, invvar, seed=392, scale=0.05)
    lf.write_db(d / "ecV.mat", V0)
    lf.write_db(d / "ecU0.mat", U0)
    ndx = [["utt0", "utt1"], ["utt2", "utt3"], ["utt4", "utt5"]]
    lf.write_lines(d / "ec.ndx", ndx)
    lf.write_cfg(d / "ec.cfg", **world["common"], ndxFilename=str(d / "ec.ndx"), inputWorldFilename="wld",
                 channelCompensation="JFA", eigenVoiceNumber=Rv, eigenVoiceMatrix="ecV", eigenChannelNumber=Ru,
                 eigenChannelMatrix="ecU_out", loadInitChannelMatrix="true", initEigenChannelMatrix="ecU0",
                 nullOrderStatSpeaker="N_ec", firstOrderStatSpeaker="FX_ec", nullOrderStatSession="Nh_ec",
                 firstOrderStatSession="FXh_ec", nbIt=2)
    _run("EigenChannel", d / "ec.cfg")
    ow = oracle.gmm(world["w"], world["mean"], world["cov"])
    mean = world["mean"].reshape(-1)
    n_sess = sum(len(l) for l in ndx)
    N, F = np.zeros((len(ndx), C)), np.zeros((len(ndx), C * D))
    Nh, Fh, spk_of = np.zeros((n_sess, C)), np.zeros((n_sess, C * D)), []
    h = 0
    for row, line in enumerate(ndx):
        for u in line:
            X = np.ascontiguousarray(world["utts"][u][_selected(u, world["utts"][u])])
            n1, f1 = oracle.bwstats(ow, X, np.zeros(len(X), dtype=np.int32), 1)
            Nh[h], Fh[h] = n1[0], f1[0]
            N[row] += n1[0]
            F[row] += f1[0]
            spk_of.append(row)
            h += 1
    Y = oracle.tv_ivectors(N, oracle.tv_subtract_m(N, F, mean), V0, invvar, oracle.tv_tett(V0, invvar, C, D))
    sup = mean[None, :] + Y @ V0                                   # M + V y per speaker
    Fc = Fh - np.repeat(Nh, D, axis=1) * sup[np.array(spk_of)]
    Ur = U0.copy()
    for it in range(2):
        _, A, Cmx, _, _, _ = oracle.tv_estep(Nh, Fc, Ur, invvar, oracle.tv_tett(Ur, invvar, C, D))
        Ur = oracle.tv_mstep(A, Cmx, C, D)
    got = lf.read_db(d / "ecU_out.mat")
    assert np.abs(got - Ur).max() < 1e-3 * np.abs(Ur).max()


def test_eigenchannel_lfa_cli(world, oracle):
    """EigenChannel --eigenChannelMode LFA (EigenChannel.cpp:178-290): D = sqrt(Sigma / tau), and per iteration the
    session statistics are centred by M + V y + D z (z from the previous iteration), x and the U accumulators come
    from them, z = tau / (tau + N) D Sigma^-1 (F_X - sum_h N_h o (M + U x_h)) on the speaker statistics."""
    d, C, D, Rv, Ru, tau = world["dir"], world["C"], world["D"], 4, 3, 14
    invvar = (1.0 / world["cov"]).reshape(-1)
    V0 = synth.make_T(Rv, C, D, invvar, seed=391, scale=0.05)
    U0 = synth.make_T(Ru, C, D, invvar, seed=392, scale=0.05)
    lf.write_db(d / "lfV.mat", V0)
    lf.write_db(d / "lfU0.mat", U0)
    ndx = [["utt0", "utt1"], ["utt2", "utt3"], ["utt4", "utt5"]]
    lf.write_lines(d / "lf.ndx", ndx)
    lf.write_cfg(d / "lf.cfg", **world["common"], ndxFilename=str(d / "lf.ndx"), inputWorldFilename="wld",
                 eigenChannelMode="LFA", eigenVoiceNumber=Rv, eigenVoiceMatrix="lfV", eigenChannelNumber=Ru,
                 eigenChannelMatrix="lfU_out", loadInitChannelMatrix="true", initEigenChannelMatrix="lfU0",
                 nullOrderStatSpeaker="N_lf", firstOrderStatSpeaker="FX_lf", nullOrderStatSession="Nh_lf",
                 firstOrderStatSession="FXh_lf", nbIt=3, regulationFactor=tau)
    _run("EigenChannel", d / "lf.cfg")
    ow = oracle.gmm(world["w"], world["mean"], world["cov"])
    mean = world["mean"].reshape(-1)
    n_sess = sum(len(l) for l in ndx)
    N, F = np.zeros((len(ndx), C)), np.zeros((len(ndx), C * D))
    Nh, Fh, spk_of = np.zeros((n_sess, C)), np.zeros((n_sess, C * D)), []
    h = 0
    for row, line in enumerate(ndx):
        for u in line:
            X = np.ascontiguousarray(world["utts"][u][_selected(u, world["utts"][u])])
            n1, f1 = oracle.bwstats(ow, X, np.zeros(len(X), dtype=np.int32), 1)
            Nh[h], Fh[h] = n1[0], f1[0]
            N[row] += n1[0]
            F[row] += f1[0]
            spk_of.append(row)
            h += 1
    spk_of = np.array(spk_of)
    Y = oracle.tv_ivectors(N, oracle.tv_subtract_m(N, F, mean), V0, invvar, oracle.tv_tett(V0, invvar, C, D))
    VY = Y @ V0
    Dm = np.sqrt(1.0 / (invvar * tau))
    Z = np.zeros((len(ndx), C * D))
    Ur = U0.copy()
    NhR, NR = np.repeat(Nh, D, axis=1), np.repeat(N, D, axis=1)
    for it in range(3):
        Fc = Fh - NhR * (mean[None, :] + VY[spk_of] + Dm[None, :] * Z[spk_of])
        Xf, A, Cmx, _, _, _ = oracle.tv_estep(Nh, Fc, Ur, invvar, oracle.tv_tett(Ur, invvar, C, D))
        Fs = F.copy()
        np.subtract.at(Fs, spk_of, NhR * (mean[None, :] + Xf @ Ur))
        Z = tau / (tau + NR) * Dm[None, :] * invvar[None, :] * Fs
        Ur = oracle.tv_mstep(A, Cmx, C, D)
    assert np.abs(Z).max() > 0                                    # the z feedback is exercised from iteration 1 on
    got = lf.read_db(d / "lfU_out.mat")
    assert np.abs(got - Ur).max() < 1e-3 * np.abs(Ur).max()
    # ... and it matters: the JFA-mode result on the same data is a different matrix
    lf.write_cfg(d / "lfj.cfg", **world["common"], ndxFilename=str(d / "lf.ndx"), inputWorldFilename="wld",
                 eigenChannelMode="JFA", eigenVoiceNumber=Rv, eigenVoiceMatrix="lfV", eigenChannelNumber=Ru,
                 eigenChannelMatrix="lfU_jfa", loadInitChannelMatrix="true", initEigenChannelMatrix="lfU0",
                 nullOrderStatSpeaker="N_lf", firstOrderStatSpeaker="FX_lf", nullOrderStatSession="Nh_lf",
                 firstOrderStatSession="FXh_lf", nbIt=3, loadAccs="true")
    _run("EigenChannel", d / "lfj.cfg")
    assert np.abs(lf.read_db(d / "lfU_jfa.mat") - Ur).max() > 1e-2 * np.abs(Ur).max()


def test_estimate_d_matrix_cli(world, oracle):
    """EstimateDMatrix (EstimateDMatrix.cpp:99-210): y with V, x with U, then the diagonal update estimateZandD
    (AccumulateJFAStat.cpp:3480-3515) on F' = F_X - N o (M + V y) - sum_h N_h o (U x_h); MAP initialisation."""
    d, C, D, Rv, Ru = world["dir"], world["C"], world["D"], 4, 3
    invvar = (1.0 / world["cov"]).reshape(-1)
    V0 = synth.make_T(Rv, C, D, invvar, seed=491, scale=0.05)
    U0 = synth.make_T(Ru, C, D, invvar, seed=492, scale=0.05)
    lf.write_db(d / "edV.mat", V0)
    lf.write_db(d / "edU.mat", U0)
    ndx = [["utt0", "utt1"], ["utt2", "utt3"], ["utt4", "utt5"]]
    lf.write_lines(d / "ed.ndx", ndx)
    reg = 14.0
    lf.write_cfg(d / "ed.cfg", **world["common"], ndxFilename=str(d / "ed.ndx"), inputWorldFilename="wld",
                 eigenVoiceNumber=Rv, eigenVoiceMatrix="edV", eigenChannelNumber=Ru, eigenChannelMatrix="edU",
                 DMatrix="edD", regulationFactor=reg, saveAllDMatrices="false", nullOrderStatSpeaker="N_ed",
                 firstOrderStatSpeaker="FX_ed", nullOrderStatSession="Nh_ed", firstOrderStatSession="FXh_ed", nbIt=3)
    _run("EstimateDMatrix", d / "ed.cfg")
    ow = oracle.gmm(world["w"], world["mean"], world["cov"])
    mean = world["mean"].reshape(-1)
    n_sess = sum(len(l) for l in ndx)
    N, F = np.zeros((len(ndx), C)), np.zeros((len(ndx), C * D))
    Nh, Fh, spk_of = np.zeros((n_sess, C)), np.zeros((n_sess, C * D)), []
    h = 0
    for row, line in enumerate(ndx):
        for u in line:
            X = np.ascontiguousarray(world["utts"][u][_selected(u, world["utts"][u])])
            n1, f1 = oracle.bwstats(ow, X, np.zeros(len(X), dtype=np.int32), 1)
            Nh[h], Fh[h] = n1[0], f1[0]
            N[row] += n1[0]
            F[row] += f1[0]
            spk_of.append(row)
            h += 1
    spk_of = np.array(spk_of)
    Y = oracle.tv_ivectors(N, oracle.tv_subtract_m(N, F, mean), V0, invvar, oracle.tv_tett(V0, invvar, C, D))
    sup = mean[None, :] + Y @ V0
    Fhc = Fh - np.repeat(Nh, D, axis=1) * sup[spk_of]
    Xs = oracle.tv_ivectors(Nh, Fhc, U0, invvar, oracle.tv_tett(U0, invvar, C, D))
    Fp = F - np.repeat(N, D, axis=1) * sup
    UX = np.repeat(Nh, D, axis=1) * (Xs @ U0)
    for hh in range(n_sess):
        Fp[spk_of[hh]] -= UX[hh]
    Dm = np.sqrt(1.0 / (invvar * reg))
    Nr = np.repeat(N, D, axis=1)
    for it in range(3):
        L = 1.0 + Nr * invvar * Dm ** 2
        Z = Fp * invvar * Dm / L
        Dm = (Z * Fp).sum(0) / ((1.0 / L + Z * Z) * Nr).sum(0)
    got = lf.read_db(d / "edD.mat")
    assert got.shape == (1, C * D) and np.abs(got[0] - Dm).max() < 1e-3 * np.abs(Dm).max()


def test_compute_test_jfa_cli(world, oracle):
    """ComputeTest --channelCompensation JFA (ComputeTest.cpp:228-370 DotProduct, :376-572 FrameByFrame): channel
    factor x of every test segment with U (y = z = 0), then either the compensated, occupation-normalised first-order
    statistics against the client supervector, or U x removed from the frames (posteriors under M + U x) followed by
    the ordinary top-K LLR against the world."""
    d, C, D, Ru = world["dir"], world["C"], world["D"], 3
    invvar = (1.0 / world["cov"]).reshape(-1)
    mean = world["mean"].reshape(-1)
    U = synth.make_T(Ru, C, D, invvar, seed=491, scale=1.0) * np.sqrt(world["cov"]).reshape(-1) * 0.4
    lf.write_db(d / "jtU.mat", U)
    rng = np.random.default_rng(492)
    # test segments WITH a channel offset: frames drawn around M + U x_true (no label file: every frame selected)
    utts = {}
    for i in range(3):
        x_true = rng.standard_normal(Ru)
        utts[f"jfa{i}"] = synth.make_frames(world["w"], world["mean"] + (x_true @ U).reshape(C, D), world["cov"] * 1.5,
                                            400 + 50 * i, seed=480 + i)
        lf.write_spro4(d / f"jfa{i}.prm", utts[f"jfa{i}"])
    clients, sups = {}, {}
    for k in range(2):
        p_ = synth.perturb_ubm(world["w"], world["mean"], world["cov"], seed=493 + k, frac=0.4, scale=0.5)
        clients[f"jspk{k}"] = p_
        lf.write_raw_gmm(d / f"jspk{k}.gmm", *p_)
        sups[f"jspk{k}"] = rng.standard_normal(C * D)
        lf.write_db(d / f"jspk{k}.sv", sups[f"jspk{k}"][None, :])
    ndx = [["jfa0", "jspk0", "jspk1"], ["jfa1", "jspk1"], ["jfa2", "jspk0"]]
    lf.write_lines(d / "jt.ndx", ndx)
    lf.write_cfg(d / "jt.cfg", **world["common"], ndxFilename=str(d / "jt.ndx"), inputWorldFilename="wld",
                 outputFilename=str(d / "jt_dot.res"), gender="M", topDistribsCount=5, computeLLKWithTopDistribs="COMPLETE",
                 channelCompensation="JFA", eigenChannelMatrix="jtU", eigenChannelNumber=Ru,
                 loadVectorFilesPath=str(d), vectorFilesExtension=".sv")
    ow = oracle.gmm(world["w"], world["mean"], world["cov"])
    tett = oracle.tv_tett(U, invvar, C, D)
    ref_dot, ref_llr, ref_lfa, ref_lfa_cms = [], [], [], []
    tau = 14
    Dm = np.sqrt(1.0 / (invvar * tau))
    for line in ndx:
        X = np.ascontiguousarray(utts[line[0]], dtype=np.float32)
        n1, f1 = oracle.bwstats(ow, X, np.zeros(len(X), dtype=np.int32), 1)
        x = oracle.tv_ivectors(n1, oracle.tv_subtract_m(n1, f1, mean), U, invvar, tett)
        ux = (x @ U)[0]
        fx = (f1[0] - np.repeat(n1[0], D) * (mean + ux)) / n1[0].sum()
        Xc = oracle.jfa_normalize_features(oracle.gmm(world["w"], world["mean"] + ux.reshape(C, D), world["cov"]), ux, X,
                                           [(0, len(X))])
        assert np.abs(Xc - X).max() > 0.1 * np.sqrt(world["cov"]).mean()    # the compensation moves the frames
        llk_w, idx, _, rest, _ = oracle.llk_determine_top(ow, Xc, 5, True)
        for cl in line[1:]:
            ref_dot.append((cl, line[0], float(sups[cl] @ fx)))
            llk_c = oracle.llk_use_top(oracle.gmm(*clients[cl]), Xc, idx, rest, True)
            ref_llr.append((cl, line[0], float(llk_c.mean() - llk_w.mean())))
        # LFA (ComputeTest.cpp:574-762): the session model also carries D z, z = tau / (tau + N) D Sigma^-1 (F - N o M)
        nrep = np.repeat(n1[0], D)
        z = tau / (tau + nrep) * Dm * invvar * (f1[0] - nrep * mean)
        Xl = oracle.jfa_normalize_features(
            oracle.gmm(world["w"], world["mean"] + (ux + Dm * z).reshape(C, D), world["cov"]), ux, X, [(0, len(X))])
        assert np.abs(Xl - Xc).max() > 1e-3
        Xs = Xl.astype(np.float64)
        Xs = ((Xs - Xs.mean(0)) / Xs.std(0)).astype(np.float32)          # cms(): zero mean, unit deviation
        for Xv, out in ((Xl, ref_lfa), (Xs, ref_lfa_cms)):
            llk_w, idx, _, rest, _ = oracle.llk_determine_top(ow, Xv, 5, True)
            for cl in line[1:]:
                llk_c = oracle.llk_use_top(oracle.gmm(*clients[cl]), Xv, idx, rest, True)
                out.append((cl, line[0], float(llk_c.mean() - llk_w.mean())))
    _run("ComputeTest", d / "jt.cfg")                       # scoring defaults to DotProduct
    got = [l.split() for l in open(d / "jt_dot.res")]
    assert [(l[1], l[3]) for l in got] == [(r[0], r[1]) for r in ref_dot]
    scale = max(abs(r[2]) for r in ref_dot)
    for l, r in zip(got, ref_dot):
        assert l[0] == "M" and abs(float(l[4]) - r[2]) < 1e-4 * scale, (l, r)
    _run("ComputeTest", d / "jt.cfg", scoring="FrameByFrame", outputFilename=str(d / "jt_fbf.res"))
    got = [l.split() for l in open(d / "jt_fbf.res")]
    assert [(l[1], l[3]) for l in got] == [(r[0], r[1]) for r in ref_llr]
    for l, r in zip(got, ref_llr):
        assert abs(float(l[4]) - r[2]) < 2e-4 and int(l[2]) == int(r[2] > 0), (l, r)
    for ref, over in ((ref_lfa, {}), (ref_lfa_cms, dict(cms="true"))):
        _run("ComputeTest", d / "jt.cfg", channelCompensation="LFA", regulationFactor=tau,
             outputFilename=str(d / "jt_lfa.res"), **over)
        got = [l.split() for l in open(d / "jt_lfa.res")]
        assert [(l[1], l[3]) for l in got] == [(r[0], r[1]) for r in ref]
        for l, r in zip(got, ref):
            assert abs(float(l[4]) - r[2]) < 2e-4 * max(1.0, abs(r[2])), (l, r)
    # without an eigenchannel matrix x = 0: FrameByFrame is the plain ComputeTest
    lf.write_cfg(d / "jt0.cfg", **world["common"], ndxFilename=str(d / "jt.ndx"), inputWorldFilename="wld",
                 outputFilename=str(d / "jt0.res"), gender="M", topDistribsCount=5, computeLLKWithTopDistribs="COMPLETE",
                 channelCompensation="JFA", scoring="FrameByFrame")
    _run("ComputeTest", d / "jt0.cfg")
    _run("ComputeTest", d / "jt0.cfg", channelCompensation="none", outputFilename=str(d / "jt0_plain.res"))
    assert open(d / "jt0.res").read() == open(d / "jt0_plain.res").read()


def test_ivextractor_approximate_modes_cli(world, oracle):
    """IvExtractor --mode ubmWeight / eigenDecomposition (IvExtractor.cpp:151-363), both computing the
    approximation parameters on the fly and loading the ones TotalVariability wrote with
    approximationMode (TotalVariability.cpp:181-241)."""
    d, C, D, R = world["dir"], world["C"], world["D"], 5
    invvar = (1.0 / world["cov"]).reshape(-1)
    T = synth.make_T(R, C, D, invvar, seed=92, scale=0.05)
    lf.write_db(d / "TVa.mat", T)
    ids = [["idA", "utt0", "utt1"], ["idB", "utt2"], ["idC", "utt3", "utt4"]]
    lf.write_lines(d / "ids_a.ndx", ids)
    ow = oracle.gmm(world["w"], world["mean"], world["cov"])
    N = np.zeros((3, C))
    F = np.zeros((3, C * D))
    for row, line in enumerate(ids):
        for u in line[1:]:
            X = np.ascontiguousarray(world["utts"][u][_selected(u, world["utts"][u])])
            n1, f1 = oracle.bwstats(ow, X, np.zeros(len(X), dtype=np.int32), 1)
            N[row] += n1[0]
            F[row] += f1[0]
    Tn = oracle.tv_norm_t(T, invvar)
    Fn = oracle.tv_norm_statistics(N, F, world["mean"].reshape(-1), invvar)
    Wcov = oracle.tv_weighted_cov(Tn, world["w"], C, D)
    Q, _ = oracle.eigen_sym(Wcov)
    Dm = oracle.tv_approximate_tctc(Tn, Q, C, D)
    W_ubm = oracle.tv_ivectors_ubm_weight(N, Fn, Tn, Wcov)
    W_eig = oracle.tv_ivectors_eigen(N, Fn, Tn, Dm, Q)

    def check(sub, W):
        for row, line in enumerate(ids):
            y = lf.read_db(d / sub / f"{line[0]}.y")
            assert y.shape == (1, R) and np.abs(y[0] - W[row]).max() < 1e-4 * np.abs(W).max(), sub

    base = dict(world["common"], targetIdList=str(d / "ids_a.ndx"), inputWorldFilename="wld",
                totalVariabilityNumber=R, totalVariabilityMatrix="TVa", nullOrderStatSpeaker="N_a",
                firstOrderStatSpeaker="F_a", vectorFilesExtension=".y")
    for mode, W in (("ubmWeight", W_ubm), ("eigenDecomposition", W_eig)):
        os.makedirs(d / f"iv_{mode}", exist_ok=True)
        lf.write_cfg(d / f"iv_{mode}.cfg", **base, saveVectorFilesPath=str(d / f"iv_{mode}") + "/", mode=mode)
        _run("IvExtractor", d / f"iv_{mode}.cfg")
        check(f"iv_{mode}", W)
    # TotalVariability with nbIt = 0 only writes the approximation parameters of the loaded matrix
    lf.write_lines(d / "tv_a.ndx", [l[1:] for l in ids])
    for mode in ("ubmWeight", "eigenDecomposition"):
        lf.write_cfg(d / f"tv_{mode}.cfg", **world["common"], ndxFilename=str(d / "tv_a.ndx"), inputWorldFilename="wld",
                     totalVariabilityNumber=R, totalVariabilityMatrix="TVa", loadInitTotalVariabilityMatrix="true",
                     initTotalVariabilityMatrix="TVa", nullOrderStatSpeaker="N_a", firstOrderStatSpeaker="F_a",
                     loadAccs="true", nbIt=0, approximationMode=mode)
        _run("TotalVariability", d / f"tv_{mode}.cfg")
        lf.write_db(d / "TVa.mat", T)   # nbIt = 0 re-saves the (unchanged) matrix; keep the input pristine
        assert np.allclose(lf.read_db(d / "TVa_norm.mat"), Tn, rtol=1e-12)
    assert np.allclose(lf.read_db(d / "TVa_weightedCov.mat"), Wcov, rtol=1e-9, atol=1e-12 * np.abs(Wcov).max())
    Qg, Dg = lf.read_db(d / "TVa_EigDec_Q.mat"), lf.read_db(d / "TVa_EigDec_D.mat")
    assert np.allclose(np.abs(Qg.T @ Q), np.eye(R), atol=1e-6) and np.allclose(Dg, Dm, rtol=1e-6, atol=1e-9 * Dm.max())
    for mode, W, flag in (("ubmWeight", W_ubm, "loadUbmWeightParam"), ("eigenDecomposition", W_eig, "loadEigenDecompositionParam")):
        os.makedirs(d / f"ivl_{mode}", exist_ok=True)
        lf.write_cfg(d / f"ivl_{mode}.cfg", **base, saveVectorFilesPath=str(d / f"ivl_{mode}") + "/", mode=mode,
                     loadAccs="true", **{flag: "true"})
        _run("IvExtractor", d / f"ivl_{mode}.cfg")
        check(f"ivl_{mode}", W)


def test_ivtest_plda_cli(world, oracle):
    """scoring = plda without ivNorm: pldaNativeScoring never centres the vectors (PldaTools.cpp:4489-4519;
    PldaTest::center is only reached through sphericalNuisanceNormalization), so pldaMeanVec is not applied."""
    d = world["dir"]
    F, G, Sigma, models, model_of, segments = synth.make_plda(d=20, rF=6, rG=3, sessions=[2, 2, 1, 1], n_test=7, seed=95)
    os.makedirs(d / "vec", exist_ok=True)
    mean = np.linspace(-0.5, 0.5, 20)
    names_e = [f"e{j}" for j in range(models.shape[1])]
    for j, n in enumerate(names_e):
        lf.write_db(d / "vec" / f"{n}.y", models[:, j][None])
    for j in range(segments.shape[1]):
        lf.write_db(d / "vec" / f"t{j}.y", segments[:, j][None])
    lf.write_db(d / "pF.mat", F)
    lf.write_db(d / "pG.mat", G)
    lf.write_db(d / "pS.mat", Sigma)
    lf.write_db(d / "pMean.mat", mean[None])
    enrol = [["m0", "e0", "e1"], ["m1", "e2", "e3"], ["m2", "e4"], ["m3", "e5"]]
    lf.write_lines(d / "enrol.ndx", enrol)
    trials = [[f"t{j}", "m0", "m1", "m2", "m3"] for j in range(7)]
    lf.write_lines(d / "trials.ndx", trials)
    lf.write_cfg(d / "it.cfg", **world["common"], ndxFilename=str(d / "trials.ndx"), targetIdList=str(d / "enrol.ndx"),
                 testVectorFilesPath=str(d / "vec"), loadVectorFilesExtension=".y", scoring="plda",
                 pldaEigenVoiceNumber=6, pldaEigenChannelNumber=3, iVectSize=20, pldaEigenVoiceMatrix="pF",
                 pldaEigenChannelMatrix="pG", pldaSigmaMatrix="pS", pldaMeanVec="pMean",
                 outputFilename=str(d / "it.res"), gender="M")
    _run("IvTest", d / "it.cfg")
    ref = oracle.plda_native_scoring(F, G, Sigma, models, model_of, segments)
    lines = [l.split() for l in open(d / "it.res")]
    assert len(lines) == 28
    for l in lines:
        m, s = int(l[1][1:]), int(l[3][1:])
        assert abs(float(l[4]) - ref[m, s]) < 1e-5 * max(1.0, abs(ref[m, s]))


def test_ivtest_lda_scatter_matrices_cli(world, oracle):
    """ldaMode scatterMatrices (PldaTools.cpp:1389, computeScatterMatUnThreaded :1607-1640), restated AS WRITTEN:
    SB = unnormalised scatter of the speaker means; SW = scatter of the FIRST n_last sessions around their own
    speakers' means / n_last (n_last = sessions of the last speaker) -- the loop assigns SW per speaker and always
    walks the sessions from 0.  Only usable when n_last >= vectSize; the test set is built that way."""
    d = world["dir"]
    dim, n_spk, per = 4, 6, 6
    rng = np.random.default_rng(131)
    os.makedirs(d / "svec", exist_ok=True)
    spk_c = rng.standard_normal((dim, n_spk)) * 1.5
    lines, cols = [], []
    for c in range(n_spk):
        names = []
        for j in range(per):
            v = spk_c[:, c] + rng.standard_normal(dim)
            names.append(f"sd{c}_{j}")
            cols.append((v, c))
            lf.write_db(d / "svec" / f"{names[-1]}.y", v[None])
        lines.append(names)
    lf.write_lines(d / "sdev.ndx", lines)
    data = np.stack([c[0] for c in cols], axis=1)
    cls = np.array([c[1] for c in cols], dtype=np.int32)
    models, segments = rng.standard_normal((dim, 3)), rng.standard_normal((dim, 3))
    for j in range(3):
        lf.write_db(d / "svec" / f"sm{j}.y", models[:, j][None])
        lf.write_db(d / "svec" / f"st{j}.y", segments[:, j][None])
    lf.write_lines(d / "strials.ndx", [[f"st{j}", "sm0", "sm1", "sm2"] for j in range(3)])
    gmean, spk_means, _, _, _ = oracle.iv_cov_mat(data, cls, n_spk)
    cm = spk_means - gmean[:, None]
    SB = cm @ cm.T
    xc = data[:, :per] - spk_means[:, cls[:per]]
    SW = xc @ xc.T / per
    lda = oracle.iv_lda(SW, SB, 2)
    ref = oracle.iv_cosine(oracle.iv_rotate_left(lda, models), oracle.iv_rotate_left(lda, segments))
    lf.write_cfg(d / "sc.cfg", **world["common"], ndxFilename=str(d / "strials.ndx"), testVectorFilesPath=str(d / "svec"),
                 loadVectorFilesPath=str(d / "svec"), loadVectorFilesExtension=".y", backgroundNdxFilename=str(d / "sdev.ndx"),
                 ivNorm="true", ivNormLoadParam="false", ivNormIterationNb=0, LDA="true", ldaRank=2,   # LDA only, no EFR
                 ldaMode="scatterMatrices", ldaMatrix="ldaScatter", gender="M",
                 wccn="false", scoring="cosine", outputFilename=str(d / "sc.res"))
    _run("IvTest", d / "sc.cfg")
    out = [l.split() for l in open(d / "sc.res")]
    assert len(out) == 9
    for l in out:
        m, sg = int(l[1][2:]), int(l[3][2:])
        assert abs(float(l[4]) - ref[m, sg]) < 1e-6, l


def test_ivtest_backend_and_ivnorm_cli(world, oracle):
    """IvTest scoring = cosine (+ WCCN) / mahalanobis / 2cov with EFR + LDA normalisation estimated on
    a development list (IvTest.cpp:112-391), and the IvNorm program (IvNorm.cpp:72-128), against the
    restated PldaTools.cpp loops."""
    d = world["dir"]
    dim, n_spk = 12, 30
    rng = np.random.default_rng(97)
    sizes = rng.integers(2, 6, n_spk)
    spk = rng.standard_normal((dim, n_spk)) * 1.5
    os.makedirs(d / "bvec", exist_ok=True)
    dev_lines, cols = [], []
    for sidx in np.argsort(-sizes, kind="stable"):          # the reference sorts speakers by session count
        names = []
        for j in range(sizes[sidx]):
            v = spk[:, sidx] + rng.standard_normal(dim) + 0.4
            names.append(f"dev{sidx}_{j}")
            cols.append((v, len(dev_lines)))
            lf.write_db(d / "bvec" / f"{names[-1]}.y", v[None])
        dev_lines.append(names)
    order = rng.permutation(len(dev_lines))                 # ... whatever the order in the file
    lf.write_lines(d / "dev.ndx", [dev_lines[i] for i in order])
    data = np.stack([c[0] for c in cols], axis=1)
    cls = np.array([c[1] for c in cols], dtype=np.int32)
    models = rng.standard_normal((dim, 4)) + 0.4
    segments = rng.standard_normal((dim, 6)) + 0.4
    for j in range(4):
        lf.write_db(d / "bvec" / f"bm{j}.y", models[:, j][None])
    for j in range(6):
        lf.write_db(d / "bvec" / f"bt{j}.y", segments[:, j][None])
    trials = [[f"bt{j}"] + [f"bm{m}" for m in range(4) if (m + j) % 3 != 0] for j in range(6)]
    lf.write_lines(d / "btrials.ndx", trials)

    # the reference chain: EFR iteration (center, whiten, length-norm) then LDA, estimated on the dev set
    mean, _, S, _, _ = oracle.iv_cov_mat(data, cls, n_spk)
    E = oracle.iv_efr_matrix(S)
    norm = lambda X: oracle.iv_length_norm(oracle.iv_rotate_left(E, oracle.iv_center(X, mean)))
    dev1 = norm(data)
    _, _, _, W1, B1 = oracle.iv_cov_mat(dev1, cls, n_spk)
    lda = oracle.iv_lda(W1, B1, 8)
    dev2 = oracle.iv_rotate_left(lda, dev1)
    m2, s2 = oracle.iv_rotate_left(lda, norm(models)), oracle.iv_rotate_left(lda, norm(segments))
    _, _, _, W2, B2 = oracle.iv_cov_mat(dev2, cls, n_spk)
    wccn = oracle.iv_wccn_chol(dev2, cls, n_spk)
    refs = {"cosine": oracle.iv_cosine(oracle.iv_rotate_left(wccn, m2), oracle.iv_rotate_left(wccn, s2)),
            "mahalanobis": oracle.iv_mahalanobis(m2, s2, oracle.invert(W2)),
            "2cov": oracle.iv_two_cov(m2, s2, W2, B2)}
    base = dict(world["common"], ndxFilename=str(d / "btrials.ndx"), testVectorFilesPath=str(d / "bvec"),
                loadVectorFilesPath=str(d / "bvec"), loadVectorFilesExtension=".y", backgroundNdxFilename=str(d / "dev.ndx"),
                ivNorm="true", ivNormLoadParam="false", ivNormIterationNb=1, ivNormEfrMode="EFR", LDA="true", ldaRank=8,
                ldaMatrix="ldaB", gender="M", wccn="true", loadWccnMatrix="false", loadMahalanobisMatrix="false",
                mahalanobisMatrix="MahB", load2covMatrix="false", TwoCovFilename="TwoCovB")
    for scoring, ref in refs.items():
        lf.write_cfg(d / f"bk_{scoring}.cfg", **base, scoring=scoring, outputFilename=str(d / f"bk_{scoring}.res"))
        _run("IvTest", d / f"bk_{scoring}.cfg")
        lines = [l.split() for l in open(d / f"bk_{scoring}.res")]
        assert len(lines) == sum(len(t) - 1 for t in trials)
        for l in lines:
            m, sg = int(l[1][2:]), int(l[3][2:])
            assert abs(float(l[4]) - ref[m, sg]) < 1e-5 * max(1.0, np.abs(ref).max()), (scoring, l)
    assert np.allclose(lf.read_db(d / "EFR_ivNormEfrMatrix_it0.mat"), E, atol=1e-7 * np.abs(E).max())
    assert np.allclose(lf.read_db(d / "EFR_ivNormEfrMean_it0.mat")[0], mean)
    # second run loading every parameter the first one saved: same scores
    lf.write_cfg(d / "bk_load.cfg", **dict(base, ivNormLoadParam="true", loadMahalanobisMatrix="true"),
                 scoring="mahalanobis", outputFilename=str(d / "bk_load.res"))
    _run("IvTest", d / "bk_load.cfg")
    assert open(d / "bk_load.res").read() == open(d / "bk_mahalanobis.res").read()
    # scoring = plda on top of the same normalisation (IvTest.cpp:301-318 runs before EVERY scoring mode):
    # the PLDA model lives in the EFR + LDA space (ldaRank = 8 dimensions)
    Fp, Gp, Sp, _, _, _ = synth.make_plda(d=8, rF=3, rG=2, sessions=[1], n_test=1, seed=99)
    lf.write_db(d / "npF.mat", Fp)
    lf.write_db(d / "npG.mat", Gp)
    lf.write_db(d / "npS.mat", Sp)
    lf.write_cfg(d / "bk_plda.cfg", **dict(base, ivNormLoadParam="true"), scoring="plda", pldaLoadModel="true",
                 pldaEigenVoiceNumber=3, pldaEigenChannelNumber=2, iVectSize=dim, pldaEigenVoiceMatrix="npF",
                 pldaEigenChannelMatrix="npG", pldaSigmaMatrix="npS", outputFilename=str(d / "bk_plda.res"))
    _run("IvTest", d / "bk_plda.cfg")
    ref = oracle.plda_native_scoring(Fp, Gp, Sp, m2, np.arange(4, dtype=np.int32), s2)
    lines = [l.split() for l in open(d / "bk_plda.res")]
    assert len(lines) == sum(len(t) - 1 for t in trials)
    for l in lines:
        m, sg = int(l[1][2:]), int(l[3][2:])
        assert abs(float(l[4]) - ref[m, sg]) < 1e-5 * max(1.0, np.abs(ref).max()), ("plda", l)
    # a model whose dimension is not that of the normalised vectors is refused, not read out of bounds
    lf.write_db(d / "npFbad.mat", np.zeros((dim, 3)))
    lf.write_cfg(d / "bk_plda_bad.cfg", **dict(base, ivNormLoadParam="true"), scoring="plda", pldaLoadModel="true",
                 pldaEigenVoiceNumber=3, pldaEigenChannelNumber=0, iVectSize=dim, pldaEigenVoiceMatrix="npFbad",
                 pldaSigmaMatrix="npS", outputFilename=str(d / "bk_plda_bad.res"))
    bad = subprocess.run([os.path.join(BIN, "IvTest"), "--config", str(d / "bk_plda_bad.cfg")], capture_output=True,
                         text=True, timeout=600)
    assert "does not match" in bad.stdout
    assert not os.path.exists(d / "bk_plda_bad.res") or not open(d / "bk_plda_bad.res").read().strip()
    # IvNorm: normalise a plain list of vectors with the saved parameters
    lf.write_lines(d / "bvlist.lst", [[f"bt{j}"] for j in range(6)])
    os.makedirs(d / "bnorm", exist_ok=True)
    lf.write_cfg(d / "ivnorm.cfg", **dict(base, ivNormLoadParam="true"), inputVectorFilename=str(d / "bvlist.lst"),
                 saveVectorFilesPath=str(d / "bnorm"), vectorFilesExtension=".y")
    _run("IvNorm", d / "ivnorm.cfg")
    for j in range(6):
        assert np.allclose(lf.read_db(d / "bnorm" / f"bt{j}.y")[0], s2[:, j], atol=1e-7)


def test_plda_training_cli(world, oracle):
    """The PLDA program (PLDA.cpp:73-101): load a development list + initial matrices, centre, three EM
    iterations, saveModel -- against the restated PldaModel::em_iteration chain; then IvTest scores with
    the trained model files."""
    d = world["dir"]
    dim, n_spk, rF, rG = 10, 24, 4, 2
    rng = np.random.default_rng(101)
    sizes = rng.integers(2, 5, n_spk)
    os.makedirs(d / "pvec", exist_ok=True)
    lines, cols = [], []
    for sidx in np.argsort(-sizes, kind="stable"):
        center = rng.standard_normal(dim) * 1.5
        names = []
        for j in range(sizes[sidx]):
            v = center + rng.standard_normal(dim) + 0.3
            names.append(f"p{sidx}_{j}")
            cols.append((v, len(lines)))
            lf.write_db(d / "pvec" / f"{names[-1]}.y", v[None])
        lines.append(names)
    lf.write_lines(d / "pdev.ndx", lines)
    data = np.stack([c[0] for c in cols], axis=1)
    cls = np.array([c[1] for c in cols], dtype=np.int32)
    F0, G0 = rng.standard_normal((dim, rF)), rng.standard_normal((dim, rG))
    S0 = np.cov(data, bias=True) + 0.05 * np.eye(dim)
    lf.write_db(d / "pF0.mat", F0)
    lf.write_db(d / "pG0.mat", G0)
    lf.write_db(d / "pS0.mat", S0)
    lf.write_db(d / "pM0.mat", np.zeros((dim, 1)))
    cfg = dict(world["common"], backgroundNdxFilename=str(d / "pdev.ndx"), loadVectorFilesPath=str(d / "pvec"),
               loadVectorFilesExtension=".y", pldaEigenVoiceNumber=rF, pldaEigenChannelNumber=rG, pldaNbIt=3,
               pldaLoadInitMatrices="true", pldaEigenVoiceMatrixInit="pF0", pldaEigenChannelMatrixInit="pG0",
               pldaSigmaMatrixInit="pS0", pldaMeanVecInit="pM0", pldaEigenVoiceMatrix="tF", pldaEigenChannelMatrix="tG",
               pldaSigmaMatrix="tS", pldaMeanVec="tMean", pldaMinDivMean="tDelta")
    lf.write_cfg(d / "plda_train.cfg", **cfg)
    _run("PLDA", d / "plda_train.cfg")
    mean = data.mean(1)
    st = (data - mean[:, None], F0, G0, S0, np.zeros(dim))
    for it in range(3):
        st = oracle.plda_em_iteration(st[0], cls, n_spk, *st[1:])
    for name, ref in (("tF", st[1]), ("tG", st[2]), ("tS", st[3])):
        got = lf.read_db(d / f"{name}.mat")
        assert got.shape == ref.shape and np.abs(got - ref).max() < 1e-6 * np.abs(ref).max(), name
    assert np.allclose(lf.read_db(d / "tMean.mat")[:, 0], mean)
    assert np.allclose(lf.read_db(d / "tDelta.mat")[:, 0], st[4], atol=1e-8)
    # score two development speakers against each other with the trained model
    lf.write_lines(d / "ptrials.ndx", [[lines[0][0], lines[0][1], lines[1][0]], [lines[1][1], lines[0][1], lines[1][0]]])
    lf.write_cfg(d / "plda_test.cfg", **dict(cfg, ndxFilename=str(d / "ptrials.ndx"), testVectorFilesPath=str(d / "pvec"),
                                           scoring="plda", iVectSize=dim, outputFilename=str(d / "plda_test.res"), gender="M"))
    _run("IvTest", d / "plda_test.cfg")
    sc = {(l.split()[1], l.split()[3]): float(l.split()[4]) for l in open(d / "plda_test.res")}
    assert len(sc) == 4
    # target trials (same speaker) outscore the non-target ones
    assert sc[(lines[0][1], lines[0][0])] > sc[(lines[1][0], lines[0][0])]
    assert sc[(lines[1][0], lines[1][1])] > sc[(lines[0][1], lines[1][1])]


def test_train_target_cli(world, oracle):
    """TrainTarget with MAPOccDep (mean + weight adaptation, 2 iterations) against the oracle's EM
    statistics + the numpy restatement of computeMAPOccDep (TrainTools.cpp:445-489, 871-904)."""
    from oracle import np_oracle
    d = world["dir"]
    lf.write_lines(d / "target.ndx", [["clientA", "utt1", "utt2"], ["clientB", "utt4"]])
    lf.write_cfg(d / "tt.cfg", **world["common"], targetIdList=str(d / "target.ndx"), inputWorldFilename="wld",
                 MAPAlgo="MAPOccDep", meanAdapt="true", weightAdapt="true", MAPRegFactorMean=14.0,
                 MAPRegFactorWeight=10.0, nbTrainIt=2, baggedFrameProbability=1.0)
    _run("TrainTarget", d / "tt.cfg")
    for cid, files in (("clientA", ["utt1", "utt2"]), ("clientB", ["utt4"])):
        X = np.ascontiguousarray(np.concatenate([world["utts"][u][_selected(u, world["utts"][u])] for u in files]))
        w0, m0, c0 = world["w"], world["mean"], world["cov"]
        w, m, c = w0, m0, c0
        for it in range(2):
            g = oracle.gmm(w, m, c)
            _, n, occ, m1, m2 = oracle.em_accumulate(g, X)
            w_ml, m_ml, c_ml = oracle.em_get(g, occ, m1, m2)
            w, m, c = np_oracle.map_occ_dep(w0, m0, c0, w_ml, m_ml, c_ml, n, r_mean=14.0, r_weight=10.0)
        gw, gm, gc = lf.read_raw_gmm(d / f"{cid}.gmm")
        assert np.allclose(gw, w, rtol=1e-3, atol=1e-7)
        assert np.abs(gm - m).max() < 1e-4 * np.abs(m).max()
        assert np.allclose(gc, c, rtol=1e-9)
    # MAPConst / MAPConst2 (TrainTools.cpp:355-419): mean-only interpolation with a constant a priori weight (with and
    # without the component weights); weights and variances stay the world's
    for algo in ("MAPConst", "MAPConst2"):
        lf.write_cfg(d / f"tt_{algo}.cfg", **world["common"], targetIdList=str(d / "target.ndx"),
                     inputWorldFilename="wld", MAPAlgo=algo, meanAdapt="true", MAPAlphaMean=0.75, nbTrainIt=2,
                     baggedFrameProbability=1.0)
        _run("TrainTarget", d / f"tt_{algo}.cfg", saveMixtureFileExtension=f".{algo}.gmm")
        for cid, files in (("clientA", ["utt1", "utt2"]), ("clientB", ["utt4"])):
            X = np.ascontiguousarray(np.concatenate([world["utts"][u][_selected(u, world["utts"][u])] for u in files]))
            w0, m0, c0 = world["w"], world["mean"], world["cov"]
            m = m0
            for it in range(2):
                g = oracle.gmm(w0, m, c0)
                _, n, occ, m1, m2 = oracle.em_accumulate(g, X)
                w_ml, m_ml, _ = oracle.em_get(g, occ, m1, m2)
                if algo == "MAPConst":
                    m = 0.75 * m0 + 0.25 * m_ml
                else:
                    m = (0.75 * w0[:, None] * m0 + 0.25 * w_ml[:, None] * m_ml) / (0.75 * w0 + 0.25 * w_ml)[:, None]
            gw, gm, gc = lf.read_raw_gmm(d / f"{cid}.{algo}.gmm")
            assert np.allclose(gw, w0, rtol=1e-12) and np.allclose(gc, c0, rtol=1e-12)
            assert np.abs(gm - m).max() < 1e-4 * np.abs(m).max()
            assert np.abs(gm - m0).max() > 1e-3 * np.abs(m0).max()


def test_train_target_jfa_cli(world, oracle):
    """TrainTarget --channelCompensation JFA (TrainTarget.cpp:393-617): joint [y; x] with the stacked [V; U] on the
    pooled statistics of a client's files, z with D on the statistics minus N o (M + V y + U x); outputs the client
    model M + V y + D z, the supervector Sigma^-1 (V y + D z) and, on request, x / y / z."""
    d, C, D, Rv, Ru = world["dir"], world["C"], world["D"], 4, 2
    invvar = (1.0 / world["cov"]).reshape(-1)
    mean = world["mean"].reshape(-1)
    sd = np.sqrt(world["cov"]).reshape(-1)
    V = synth.make_T(Rv, C, D, invvar, seed=591, scale=1.0) * sd * 0.3
    U = synth.make_T(Ru, C, D, invvar, seed=592, scale=1.0) * sd * 0.3
    Dm = 0.2 * sd * (1.0 + 0.5 * np.random.default_rng(593).random(C * D))
    lf.write_db(d / "ttV.mat", V)
    lf.write_db(d / "ttU.mat", U)
    lf.write_db(d / "ttD.mat", Dm[None, :])
    ids = [["jclA", "utt0", "utt2"], ["jclB", "utt3"]]
    lf.write_lines(d / "ttj.ndx", ids)
    os.makedirs(d / "jsv", exist_ok=True)
    lf.write_cfg(d / "ttj.cfg", **world["common"], targetIdList=str(d / "ttj.ndx"), inputWorldFilename="wld",
                 channelCompensation="JFA", eigenVoiceMatrix="ttV", eigenChannelMatrix="ttU", DMatrix="ttD",
                 saveVectorFilesPath=str(d / "jsv") + "/", vectorFilesExtension=".sv", saveY="true",
                 yExtension=".yfac")
    cwd = os.getcwd()
    os.chdir(d)                         # x / y / z land relative to the working directory, like the reference's
    try:
        _run("TrainTarget", d / "ttj.cfg")
    finally:
        os.chdir(cwd)
    ow = oracle.gmm(world["w"], world["mean"], world["cov"])
    VU = np.concatenate([V, U])
    tett = oracle.tv_tett(VU, invvar, C, D)
    for line in ids:
        X = np.ascontiguousarray(np.concatenate([world["utts"][u][_selected(u, world["utts"][u])] for u in line[1:]]))
        n1, f1 = oracle.bwstats(ow, X, np.zeros(len(X), dtype=np.int32), 1)
        yx = oracle.tv_ivectors(n1, oracle.tv_subtract_m(n1, f1, mean), VU, invvar, tett)[0]
        y = yx[:Rv]
        nrep = np.repeat(n1[0], D)
        fp = f1[0] - nrep * (mean + yx @ VU)
        z = fp * invvar * Dm / (1.0 + nrep * invvar * Dm * Dm)
        off = y @ V + Dm * z
        gw, gm, gc = lf.read_raw_gmm(d / f"{line[0]}.gmm")
        assert np.allclose(gw, world["w"], rtol=1e-12) and np.allclose(gc, world["cov"], rtol=1e-12)
        assert np.abs(off).max() > 0.02 * sd.mean()
        assert np.abs(gm.reshape(-1) - (mean + off)).max() < 1e-4 * np.abs(off).max()
        sv = lf.read_db(d / "jsv" / f"{line[0]}.sv").reshape(-1)
        assert np.abs(sv - off * invvar).max() < 1e-4 * np.abs(off * invvar).max()
        got_y = lf.read_db(d / f"{line[0]}.yfac").reshape(-1)
        assert np.abs(got_y - y).max() < 1e-4 * np.abs(yx).max()
    # LFA (TrainTarget.cpp:620-760): D = sqrt(Sigma / tau), z = tau / (tau + N) D Sigma^-1 F', the model only
    tau = 12
    _run("TrainTarget", d / "ttj.cfg", channelCompensation="LFA", regulationFactor=tau,
         mixtureFilesPath=str(d) + "/", saveMixtureFileExtension=".lfa.gmm")
    Dl = np.sqrt(1.0 / (invvar * tau))
    for line in ids:
        X = np.ascontiguousarray(np.concatenate([world["utts"][u][_selected(u, world["utts"][u])] for u in line[1:]]))
        n1, f1 = oracle.bwstats(ow, X, np.zeros(len(X), dtype=np.int32), 1)
        yx = oracle.tv_ivectors(n1, oracle.tv_subtract_m(n1, f1, mean), VU, invvar, tett)[0]
        nrep = np.repeat(n1[0], D)
        z = tau / (tau + nrep) * Dl * invvar * (f1[0] - nrep * (mean + yx @ VU))
        off = yx[:Rv] @ V + Dl * z
        _, gm, _ = lf.read_raw_gmm(d / f"{line[0]}.lfa.gmm")
        assert np.abs(gm.reshape(-1) - (mean + off)).max() < 1e-4 * np.abs(off).max()


def test_topgauss_index_file(world, oracle):
    """TopGauss::compute / write / read (TopGauss.cpp:68-200) through the host class (driven by HostSelfTest): per
    selected frame the retained top components (a fixed count, or as many as exceed a share of the frame likelihood),
    1 - sum of their weights, the likelihood outside them; the binary layout nt, nbgcnt, nbg[], idx[], snsw[], snsl[]."""
    d, K = world["dir"], 5
    lf.write_lines(d / "tg.ndx", [["utt2", "x"], ["utt3", "y"]])
    lf.write_cfg(d / "tg.cfg", **world["common"], ndxFilename=str(d / "tg.ndx"), inputWorldFilename="wld",
                 tmpPrefix=str(d / "tg_tmp"), topDistribsCount=K, computeLLKWithTopDistribs="COMPLETE",
                 nbGaussianFilesDir=str(d) + "/")
    X = np.ascontiguousarray(world["utts"]["utt2"][_selected("utt2", world["utts"]["utt2"])])
    ow = oracle.gmm(world["w"], world["mean"], world["cov"])
    llk, idx, top_lk, _, _ = oracle.llk_determine_top(ow, X, K, True)
    for top_d in (3, 0.9):
        out = _run("HostSelfTest", d / "tg.cfg", topGauss=top_d)
        r = {l.split()[0]: l.split()[1:] for l in out.strip().splitlines()}
        nt, cnt, mean_llk, same = int(r["topgauss"][0]), int(r["topgauss"][1]), float(r["topgauss"][2]), r["topgauss"][3]
        raw = open(d / "utt2.tg", "rb").read()
        assert same == "1" and nt == len(X) and tuple(np.frombuffer(raw, "<u8", 2)) == (nt, cnt)
        nbg = np.frombuffer(raw, "<u8", nt, 16)
        fidx = np.frombuffer(raw, "<u8", cnt, 16 + 8 * nt)
        snsw = np.frombuffer(raw, "<f8", nt, 16 + 8 * nt + 8 * cnt)
        snsl = np.frombuffer(raw, "<f8", nt, 16 + 16 * nt + 8 * cnt)
        assert len(raw) == 16 + 24 * nt + 8 * cnt
        lk_tot = np.exp(llk)
        if top_d >= 1:
            ref_n = np.full(nt, int(top_d))
        else:  # the count that first pushes the running sum over the share (checked before each addition, :173-179)
            cum = np.concatenate([np.zeros((nt, 1)), np.cumsum(top_lk, axis=1)], axis=1)[:, :K]
            ref_n = (cum <= top_d * lk_tot[:, None]).sum(1)
        assert np.array_equal(nbg, ref_n) and cnt == ref_n.sum()
        ref_idx = np.concatenate([idx[t, :ref_n[t]] for t in range(nt)])
        assert np.array_equal(fidx, ref_idx)
        ref_w = np.array([1.0 - world["w"][idx[t, :ref_n[t]]].sum() for t in range(nt)])
        ref_l = np.maximum(np.array([lk_tot[t] - top_lk[t, :ref_n[t]].sum() for t in range(nt)]), 1e-200)
        assert np.allclose(snsw, ref_w, rtol=1e-9, atol=1e-12)
        assert np.allclose(snsl, ref_l, rtol=1e-6, atol=1e-9 * lk_tot.max())
        assert abs(mean_llk - llk.mean()) < 2e-4


def test_compute_test_window_llr_cli(world, oracle):
    """ComputeTest --windowLLR (WindowLLR, UnsupervisedTools.cpp:92-150; ComputeTest.cpp:100, 143, 165-178): a result
    line per client for every window of windowLLRSize selected frames (shift windowLLRDec; the window runs across
    the selected segments), written in frame order before the file-level lines."""
    d, K, size, dec = world["dir"], 5, 40, 15
    clients = {}
    for k in range(2):
        p = synth.perturb_ubm(world["w"], world["mean"], world["cov"], seed=170 + k, frac=0.4, scale=0.5)
        clients[f"wspk{k}"] = p
        lf.write_raw_gmm(d / f"wspk{k}.gmm", *p)
    lf.write_lines(d / "win.ndx", [["utt1", "wspk0", "wspk1"], ["utt5", "wspk1"]])
    lf.write_cfg(d / "win.cfg", **world["common"], ndxFilename=str(d / "win.ndx"), inputWorldFilename="wld",
                 outputFilename=str(d / "win.res"), gender="F", topDistribsCount=K, computeLLKWithTopDistribs="COMPLETE",
                 windowLLR="true", windowLLRSize=size)
    ow = oracle.gmm(world["w"], world["mean"], world["cov"])

    def reference(dec_):
        ref = []
        for line in [["utt1", "wspk0", "wspk1"], ["utt5", "wspk1"]]:
            sel = _selected(line[0], world["utts"][line[0]])
            X = np.ascontiguousarray(world["utts"][line[0]][sel])
            llkw, idx, _, rest, _ = oracle.llk_determine_top(ow, X, K, True)
            llr = np.stack([oracle.llk_use_top(oracle.gmm(*clients[c]), X, idx, rest, True) - llkw for c in line[1:]], 1)
            # the reference's ring buffer, restated on the frame list: positions [lo, hi] of the current window
            lo, cnt = 0, 0
            for t in range(len(X)):
                if cnt < size:
                    cnt += 1
                else:
                    lo += dec_
                    cnt -= dec_ - 1
                if cnt == size:
                    hi = lo + cnt - 1
                    assert hi == t
                    for i, c in enumerate(line[1:]):
                        ref.append((c, line[0], sel[lo] * 0.01, sel[hi] * 0.01, llr[lo:hi + 1, i].sum() / size))
            for i, c in enumerate(line[1:]):
                ref.append((c, line[0], None, None, llr[:, i].mean()))
        return ref

    for dec_, over in ((size, {}), (dec, dict(windowLLRDec=dec))):
        _run("ComputeTest", d / "win.cfg", **over)
        got = [l.split() for l in open(d / "win.res")]
        ref = reference(dec_)
        assert len(got) == len(ref) and sum(r[2] is not None for r in ref) > 10
        for l, r in zip(got, ref):
            assert (l[1], l[3]) == (r[0], r[1]), (l, r)
            if r[2] is None:
                assert len(l) == 5 and abs(float(l[4]) - r[4]) < 2e-4
            else:
                assert len(l) == 7 and abs(float(l[4]) - r[2]) < 1e-6 and abs(float(l[5]) - r[3]) < 1e-6
                assert abs(float(l[6]) - r[4]) < 2e-4, (l, r)


def test_compute_stats_ivector_mode_cli(world, oracle):
    """ComputeJFAStats --computeStatMode ivector (ComputeJFAStatsMain.cpp:108-117, ComputeTVStats :89-103): the TVAcc
    statistics of every NDX line, saved under nullOrderStatSpeaker / firstOrderStatSpeaker."""
    d, C, D = world["dir"], world["C"], world["D"]
    ndx = [["utt0", "utt1"], ["utt3"]]
    lf.write_lines(d / "cs.ndx", ndx)
    lf.write_cfg(d / "cs.cfg", **world["common"], ndxFilename=str(d / "cs.ndx"), inputWorldFilename="wld",
                 computeStatMode="ivector", nullOrderStatSpeaker="N_cs", firstOrderStatSpeaker="FX_cs",
                 totalVariabilityNumber=4)
    _run("ComputeJFAStats", d / "cs.cfg")
    ow = oracle.gmm(world["w"], world["mean"], world["cov"])
    N, F = lf.read_db(d / "N_cs.mat"), lf.read_db(d / "FX_cs.mat")
    assert N.shape == (2, C) and F.shape == (2, C * D)
    for row, line in enumerate(ndx):
        X = np.ascontiguousarray(np.concatenate([world["utts"][u][_selected(u, world["utts"][u])] for u in line]))
        n1, f1 = oracle.bwstats(ow, X, np.zeros(len(X), dtype=np.int32), 1)
        assert np.abs(N[row] - n1[0]).max() < 1e-4 * n1[0].max()
        assert np.abs(F[row] - f1[0]).max() < 1e-4 * np.abs(f1[0]).max()


def test_train_target_init_by_client_and_nap_cli(world, oracle):
    """TrainTarget options around the MAP loop (TrainTarget.cpp:96-102, 136-157): initByClient (EM starts from the
    client's existing model, the world stays the a priori), NAP (the adapted mean supervector loses its projection on
    the channel subspace, computeNap SuperVectors.cpp:128-138), saveEmptyModel."""
    from oracle import np_oracle
    d, C, D = world["dir"], world["C"], world["D"]
    w0, m0, c0 = world["w"], world["mean"], world["cov"]
    start = synth.perturb_ubm(w0, m0, c0, seed=601, frac=1.0, scale=0.3)
    lf.write_raw_gmm(d / "nclA.gmm", *start)
    Q, _ = np.linalg.qr(np.random.default_rng(602).standard_normal((C * D, 3)))
    U = np.ascontiguousarray(Q.T)                                  # orthonormal rows [3 x C*D]
    lf.write_db(d / "nap.mat", U)
    lf.write_lines(d / "nt.ndx", [["nclA", "utt1", "utt2"]])
    lf.write_cfg(d / "nt.cfg", **world["common"], targetIdList=str(d / "nt.ndx"), inputWorldFilename="wld",
                 MAPAlgo="MAPOccDep", meanAdapt="true", MAPRegFactorMean=14.0, nbTrainIt=1, baggedFrameProbability=1.0,
                 initByClient="true", NAP=str(d / "nap.mat"))
    _run("TrainTarget", d / "nt.cfg", saveMixtureFileExtension=".nap.gmm")
    X = np.ascontiguousarray(np.concatenate([world["utts"][u][_selected(u, world["utts"][u])] for u in ("utt1", "utt2")]))
    g = oracle.gmm(*start)                                         # EM statistics under the client's own model
    _, n, occ, m1, m2 = oracle.em_accumulate(g, X)
    w_ml, m_ml, c_ml = oracle.em_get(g, occ, m1, m2)
    _, m, _ = np_oracle.map_occ_dep(w0, m0, c0, w_ml, m_ml, c_ml, n, r_mean=14.0, r_weight=None)
    v = m.reshape(-1)
    v = v - U.T @ (U @ v)
    gw, gm, gc = lf.read_raw_gmm(d / "nclA.nap.gmm")
    assert np.abs(gm.reshape(-1) - v).max() < 1e-4 * np.abs(v).max()
    assert np.abs(U @ gm.reshape(-1)).max() < 1e-9 * np.abs(v).max()      # nothing left in the channel subspace
    assert np.allclose(gw, w0, rtol=1e-12) and np.allclose(gc, c0, rtol=1e-12)
