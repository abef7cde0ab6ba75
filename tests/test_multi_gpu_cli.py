"""Two-GPU runs of the C++ host programs (one process per GPU, NCCL through the C ABI: lr_comm_init_file,
lr_allreduce_host, lr_tv_exchange_sharded, lr_allgather_host) against the single-process runs of the same
programs -- the e1 / e2 / e3 / e4 / e5 rows of SURVEY 8e behind the reference's command lines.
Skipped on a single-GPU box (run with `gpurun --gpus 2`)."""
import os
import subprocess

import numpy as np
import pytest

from lia_ral_b200 import synth
from tests import lia_files as lf
from tests.test_cli_gpu import BIN, _run, world  # noqa: F401  (the module-scoped fixture is reused)

pytestmark = pytest.mark.gpu


def _need_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")


def _run2(prog, cfg, comm_file, **over):
    """the same program as two ranks: --lrWorldSize 2 --lrRank r --lrCommFile <file>"""
    if os.path.exists(comm_file):
        os.remove(comm_file)
    procs = []
    for r in range(2):
        cmd = [os.path.join(BIN, prog), "--config", str(cfg), "--lrWorldSize", "2", "--lrRank", str(r),
               "--lrCommFile", str(comm_file)]
        for k, v in over.items():
            cmd += [f"--{k}", str(v)]
        procs.append(subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        out, _ = p.communicate(timeout=600)
        assert p.returncode == 0, out
        assert "Exception" not in out and "error" not in out.lower(), out
        outs.append(out)
    return outs


def test_train_world_two_ranks(world):
    """e2: segments sharded, ONE all-reduce of {occ, m1, m2, llk, n} per iteration; rank 0 writes the model."""
    _need_two_gpus()
    d = world["dir"]
    start = synth.perturb_ubm(world["w"], world["mean"], world["cov"], seed=181, frac=1.0, scale=0.4)
    lf.write_raw_gmm(d / "mstart.gmm", *start)
    lf.write_lines(d / "mtrain.lst", [[f"utt{i}"] for i in range(6)])
    cfg = dict(world["common"], inputFeatureFilename=str(d / "mtrain.lst"), inputWorldFilename="mstart", nbTrainIt=3,
               baggedFrameProbability=1.0, initVarianceFlooring=0.4, finalVarianceFlooring=0.2,
               initVarianceCeiling=8.0, finalVarianceCeiling=6.0)
    lf.write_cfg(d / "mtw.cfg", **cfg, outputWorldFilename="mtrained1")
    _run("TrainWorld", d / "mtw.cfg")
    _run2("TrainWorld", d / "mtw.cfg", d / "comm_tw", outputWorldFilename="mtrained2")
    a, b = lf.read_raw_gmm(d / "mtrained1.gmm"), lf.read_raw_gmm(d / "mtrained2.gmm")
    # (a different split of the segments regroups the fp32 chunk sums of the accumulate kernel: 1e-6, not 1e-12)
    for x, y in zip(a, b):
        assert np.allclose(x, y, rtol=2e-5, atol=1e-9)


def test_compute_test_two_ranks(world):
    """e4: NDX lines sharded, no collective; the concatenated parts are the single-process file."""
    _need_two_gpus()
    d = world["dir"]
    for k in range(3):
        lf.write_raw_gmm(d / f"mspk{k}.gmm", *synth.perturb_ubm(world["w"], world["mean"], world["cov"], seed=170 + k,
                                                                 frac=0.4, scale=0.5))
    lf.write_lines(d / "mtest.ndx", [["utt0", "mspk0", "mspk1"], ["utt3", "mspk2"], ["utt5", "mspk0", "mspk1", "mspk2"],
                                     ["utt1", "mspk1"], ["utt2", "mspk2", "mspk0"]])
    lf.write_cfg(d / "mct.cfg", **world["common"], ndxFilename=str(d / "mtest.ndx"), inputWorldFilename="wld",
                 outputFilename=str(d / "mct1.res"), gender="F", topDistribsCount=5, computeLLKWithTopDistribs="COMPLETE")
    _run("ComputeTest", d / "mct.cfg")
    _run2("ComputeTest", d / "mct.cfg", d / "comm_ct", outputFilename=str(d / "mct2.res"))
    assert open(d / "mct1.res").read() == open(d / "mct2.res").read()
    assert not os.path.exists(str(d / "mct2.res") + ".part0")


def test_ivextractor_and_total_variability_two_ranks(world):
    """e1 + e3: NDX lines sharded for the statistics and the i-vector solve (one file per line, nothing to
    gather); TotalVariability exchanges component-sharded (lr_tv_exchange_sharded) and every rank ends with the
    same T -- equal to the single-process T."""
    _need_two_gpus()
    d, C, D, R = world["dir"], world["C"], world["D"], 4     # C = 32 is divisible by 2: sharded M-step
    invvar = (1.0 / world["cov"]).reshape(-1)
    lf.write_db(d / "mTV.mat", synth.make_T(R, C, D, invvar, seed=191, scale=0.05))
    ids = [["mA", "utt0", "utt1"], ["mB", "utt2"], ["mC", "utt3", "utt4", "utt0"], ["mD", "utt5"], ["mE", "utt1", "utt2"]]
    lf.write_lines(d / "mids.ndx", ids)
    base = dict(world["common"], targetIdList=str(d / "mids.ndx"), inputWorldFilename="wld", totalVariabilityNumber=R,
                totalVariabilityMatrix="mTV", vectorFilesExtension=".y")
    for tag in ("1", "2"):
        os.makedirs(d / f"miv{tag}", exist_ok=True)
    lf.write_cfg(d / "miv.cfg", **base)
    _run("IvExtractor", d / "miv.cfg", saveVectorFilesPath=str(d / "miv1") + "/", nullOrderStatSpeaker="mN1",
         firstOrderStatSpeaker="mF1")
    _run2("IvExtractor", d / "miv.cfg", d / "comm_iv", saveVectorFilesPath=str(d / "miv2") + "/",
          nullOrderStatSpeaker="mN2", firstOrderStatSpeaker="mF2")
    for line in ids:
        assert np.allclose(lf.read_db(d / "miv1" / f"{line[0]}.y"), lf.read_db(d / "miv2" / f"{line[0]}.y"),
                           rtol=1e-9, atol=1e-12)
    # the gathered statistics files equal the single-process ones
    assert np.allclose(lf.read_db(d / "mN1.mat"), lf.read_db(d / "mN2.mat"), rtol=1e-12)
    assert np.allclose(lf.read_db(d / "mF1.mat"), lf.read_db(d / "mF2.mat"), rtol=1e-12)
    lf.write_lines(d / "mtv.ndx", [l[1:] for l in ids])
    tvcfg = dict(world["common"], ndxFilename=str(d / "mtv.ndx"), inputWorldFilename="wld", totalVariabilityNumber=R,
                 loadInitTotalVariabilityMatrix="true", initTotalVariabilityMatrix="mTV", nbIt=2, minDivergence="true")
    lf.write_cfg(d / "mtv.cfg", **tvcfg)
    _run("TotalVariability", d / "mtv.cfg", totalVariabilityMatrix="mTV_out1", nullOrderStatSpeaker="mNt1",
         firstOrderStatSpeaker="mFt1", meanEstimate="mMean1")
    _run2("TotalVariability", d / "mtv.cfg", d / "comm_tv", totalVariabilityMatrix="mTV_out2",
          nullOrderStatSpeaker="mNt2", firstOrderStatSpeaker="mFt2", meanEstimate="mMean2")
    T1, T2 = lf.read_db(d / "mTV_out1.mat"), lf.read_db(d / "mTV_out2.mat")
    assert np.abs(T1 - T2).max() < 1e-8 * np.abs(T1).max()
    # (the TV contractions run as 6-plane digit products, 2^-42 of row x column scale: a different sharding
    # regroups the batches, so two runs agree to ~1e-10 in the max norm, not element-wise to 1e-12)
    M1, M2 = lf.read_db(d / "mMean1.mat"), lf.read_db(d / "mMean2.mat")
    assert np.abs(M1 - M2).max() < 1e-8 * np.abs(M1).max()


def test_ivtest_plda_two_ranks(world):
    """e5: model rows sharded across the ranks, segments replicated; the gathered score file is the
    single-process file up to the split-precision scale choice of each shard."""
    _need_two_gpus()
    d = world["dir"]
    F, G, Sigma, models, model_of, segments = synth.make_plda(d=20, rF=6, rG=3, sessions=[2, 2, 1, 1, 3], n_test=7, seed=195)
    os.makedirs(d / "mvec", exist_ok=True)
    for j in range(models.shape[1]):
        lf.write_db(d / "mvec" / f"e{j}.y", models[:, j][None])
    for j in range(7):
        lf.write_db(d / "mvec" / f"t{j}.y", segments[:, j][None])
    lf.write_db(d / "mpF.mat", F)
    lf.write_db(d / "mpG.mat", G)
    lf.write_db(d / "mpS.mat", Sigma)
    lf.write_lines(d / "menrol.ndx", [["m0", "e0", "e1"], ["m1", "e2", "e3"], ["m2", "e4"], ["m3", "e5"], ["m4", "e6", "e7", "e8"]])
    lf.write_lines(d / "mtrials.ndx", [[f"t{j}", "m0", "m1", "m2", "m3", "m4"] for j in range(7)])
    lf.write_cfg(d / "mit.cfg", **world["common"], ndxFilename=str(d / "mtrials.ndx"), targetIdList=str(d / "menrol.ndx"),
                 testVectorFilesPath=str(d / "mvec"), loadVectorFilesExtension=".y", scoring="plda",
                 pldaEigenVoiceNumber=6, pldaEigenChannelNumber=3, iVectSize=20, pldaEigenVoiceMatrix="mpF",
                 pldaEigenChannelMatrix="mpG", pldaSigmaMatrix="mpS", gender="M")
    _run("IvTest", d / "mit.cfg", outputFilename=str(d / "mit1.res"))
    _run2("IvTest", d / "mit.cfg", d / "comm_it", outputFilename=str(d / "mit2.res"))
    a = [l.split() for l in open(d / "mit1.res")]
    b = [l.split() for l in open(d / "mit2.res")]
    assert len(a) == 35 and [l[:4] for l in a] == [l[:4] for l in b]
    sa, sb = np.array([float(l[4]) for l in a]), np.array([float(l[4]) for l in b])
    assert np.abs(sa - sb).max() < 2e-6 * np.abs(sa).max()
