"""Generate tests/golden/real_field_data.npz: REAL BattGP training sets produced by the reference's own data layer
(/root/reference/src/batt_data: feather cache -> segment filter -> OCV lookup -> R = (U - OCV)/I -> even sub-sampling,
batt_data.py:180-256) from its test fixtures (tests/data/cache/{3,14}.feather), together with the oracle's exact-GP results
on them at the reference's default hyper-parameters (config.py:39-43) and query grid (battgp_full.py:98, gp_runner.py:32).
Run in the build container only (the reference tree does not travel):  python tests/golden/make_real_data_golden.py"""
import os
import pathlib
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

import src.config as cfg                                             # noqa: E402  (reference code)
from src.batt_data.batt_data import BattData                         # noqa: E402
from src.batt_data.data_utils import read_cell_characteristics       # noqa: E402
from oracle import gp_oracle as orc                                  # noqa: E402


def main():
    cfg.PATH_DATA_CACHE = pathlib.Path(REF) / "tests" / "data" / "cache"
    cfg.PATH_FIELDDATA_DATA = "<not needed: cache hit>"
    ocv = read_cell_characteristics(pathlib.Path(REF) / "tests" / "data" / "ocv_linear_approx.csv")
    out = {}
    spec = orc.battgp_spec(cfg.OUTPUTSCALE_WIENER, cfg.OUTPUTSCALE_RBF, cfg.LENGTHSCALE_RBF)
    noise = float(cfg.NOISE_VARIANCE[0])
    for batt, cell, n in (("14", 1, 1000), ("14", -1, 1000), ("3", 5, 800), ("14", 3, 3000)):
        bd = BattData(batt, ocv)
        x, y = bd.generateTrainingData(cell, max_training_data=n, max_age=None)
        x = np.ascontiguousarray(np.asarray(x, dtype=np.float64))
        y = np.ascontiguousarray(np.asarray(y, dtype=np.float64).reshape(-1))
        t = np.linspace(x[0, 0], bd.age, 300)
        xq = np.column_stack([t, np.full(300, -15.0), np.full(300, 90.0), np.full(300, 25.0)])
        f = orc.fit(spec, x, y, noise)
        mean, var = orc.predict(spec, x, f, xq)
        key = f"b{batt}_c{cell if cell >= 0 else 'pack'}"
        out.update({f"{key}_x": x, f"{key}_y": y, f"{key}_xq": xq, f"{key}_mean": mean, f"{key}_var": var,
                    f"{key}_lml": np.array(f.lml), f"{key}_jitter": np.array(f.jitter)})
        print(key, x.shape, "age", bd.age, "lml", f.lml, "jitter", f.jitter, "mean[0]", mean[0], "var[0]", var[0])
    out["theta"] = np.array([noise, cfg.OUTPUTSCALE_WIENER, cfg.OUTPUTSCALE_RBF, *cfg.LENGTHSCALE_RBF])
    np.savez_compressed(os.path.join(HERE, "real_field_data.npz"), **out)
    print("wrote real_field_data.npz", {k: v.shape for k, v in out.items() if k.endswith("_x")})


def main_full(n_max=40000, block=4000):
    """BASELINE config 2's real-data variant: every filtered row of system 14, cell 1, capped at 40 000,
    with the oracle's results at FULL size (K assembled row-block by row-block to stay inside host RAM, LAPACK dpotrf in
    place: ~1-2 min on 8 cores, 13 GB).  Writes real_field_data_40k.npz."""
    import scipy.linalg as sla
    cfg.PATH_DATA_CACHE = pathlib.Path(REF) / "tests" / "data" / "cache"
    cfg.PATH_FIELDDATA_DATA = "<not needed: cache hit>"
    ocv = read_cell_characteristics(pathlib.Path(REF) / "tests" / "data" / "ocv_linear_approx.csv")
    bd = BattData("14", ocv)
    x, y = bd.generateTrainingData(1, max_training_data=n_max, max_age=None)
    x = np.ascontiguousarray(np.asarray(x, dtype=np.float64))
    y = np.ascontiguousarray(np.asarray(y, dtype=np.float64).reshape(-1))
    n = y.shape[0]
    print("rows", n, "age", bd.age, flush=True)
    spec = orc.battgp_spec(cfg.OUTPUTSCALE_WIENER, cfg.OUTPUTSCALE_RBF, cfg.LENGTHSCALE_RBF)
    noise = float(cfg.NOISE_VARIANCE[0])
    import torch
    k = np.empty((n, n))
    for b0 in range(0, n, block):
        e = min(n, b0 + block)
        k[b0:e, :] = orc.cov(spec, x[b0:e], x)
    k[np.diag_indices_from(k)] += noise
    # torch's LAPACK (scipy's OpenBLAS dpotrf crashes at this size in this container); same dpotrf algorithm, fp64
    kt = torch.from_numpy(k)
    ct, info = torch.linalg.cholesky_ex(kt)
    assert int(info) == 0, int(info)              # no jitter needed on this set
    del kt, k
    c = ct.numpy()
    yt = torch.from_numpy(y).reshape(-1, 1)
    z = torch.linalg.solve_triangular(ct, yt, upper=False)
    alpha = torch.linalg.solve_triangular(ct.T, z, upper=True).reshape(-1).numpy()
    z = z.reshape(-1).numpy()
    logdet = 2.0 * np.log(np.diag(c)).sum()
    lml = -0.5 * float(z @ z) - 0.5 * logdet - 0.5 * n * np.log(2.0 * np.pi)
    t = np.linspace(x[0, 0], bd.age, 300)
    xq = np.column_stack([t, np.full(300, -15.0), np.full(300, 90.0), np.full(300, 25.0)])
    kq = orc.cov(spec, xq, x)
    mean = kq @ alpha
    v = torch.linalg.solve_triangular(ct, torch.from_numpy(np.ascontiguousarray(kq.T)), upper=False).numpy()
    var = np.maximum(orc.cov_diag(spec, xq) - (v * v).sum(axis=0), 1e-10)
    # leading 2048 x 2048 block of L and a strided sample of alpha: direct checks of the factor at full size
    np.savez_compressed(os.path.join(HERE, "real_field_data_40k.npz"), x=x, y=y, xq=xq, mean=mean, var=var,
                        lml=np.array(lml), logdet=np.array(logdet), alpha_sample=alpha[::97].copy(),
                        l_diag=np.diag(c).copy(), l_last_row=c[-1, :].copy(),
                        theta=np.array([noise, cfg.OUTPUTSCALE_WIENER, cfg.OUTPUTSCALE_RBF, *cfg.LENGTHSCALE_RBF]))
    print("lml", lml, "logdet", logdet, "mean[:3]", mean[:3], "var[:3]", var[:3], flush=True)


if __name__ == "__main__":
    if "--full" in sys.argv:
        main_full()
    else:
        main()
