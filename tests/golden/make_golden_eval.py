"""Golden vectors for the evaluation helpers (``optimize_gauge_SO3`` / ``optimize_gauge_SE3`` /
``distance_SO3`` / ``angle``), produced by the REAL reference (``/root/reference/vican/geometry.py``)
in the build container:  ``python tests/golden/make_golden_eval.py``  ->  ``eval_helpers.npz``."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
sys.path.insert(0, "/root/reference")

from vican.geometry import SE3 as RefSE3, angle, distance_SO3, optimize_gauge_SE3, optimize_gauge_SO3  # noqa: E402

from vican_b200 import synthetic as syn  # noqa: E402


def main():
    rng = np.random.default_rng(123)
    n = 257
    # "estimates" = ground truth moved by one common gauge + noise (the situation of main.ipynb cell 9)
    Rgt = syn.random_rotations(rng, n)
    tgt = rng.normal(0.0, 5.0, (n, 3))
    Rg = syn.random_rotations(rng, 1)[0]
    tg = rng.normal(0.0, 2.0, 3)
    Rest = (Rgt @ Rg) @ syn.so3_exp(rng.normal(0.0, 0.01, (n, 3)))
    test = np.einsum("nij,j->ni", Rgt, tg) + tgt + rng.normal(0.0, 0.01, (n, 3))
    a = [RefSE3(R=Rgt[i], t=tgt[i]) for i in range(n)]
    b = [RefSE3(R=Rest[i], t=test[i]) for i in range(n)]
    G = optimize_gauge_SE3(a, b)
    Gr = optimize_gauge_SO3([x.R() for x in a], [x.R() for x in b])
    dist = np.array([distance_SO3(Rgt[i], Rest[i]) for i in range(n)])
    ang = np.array([angle(Rest[i]) for i in range(n)])
    # a few exactly-equal and 180-degree pairs (clip branch of geometry.py:150)
    Rspecial = np.stack([np.eye(3), np.diag([1.0, -1.0, -1.0]), np.diag([-1.0, -1.0, 1.0])])
    dist_special = np.array([distance_SO3(np.eye(3), r) for r in Rspecial])
    np.savez_compressed(os.path.join(HERE, "eval_helpers.npz"), Rgt=Rgt, tgt=tgt, Rest=Rest, test=test,
                        gauge_R=np.asarray(G.R(), np.float64), gauge_t=np.asarray(G.t(), np.float64).reshape(3),
                        gauge_so3=np.asarray(Gr, np.float64), dist_deg=dist, angle_deg=ang, Rspecial=Rspecial,
                        dist_special=dist_special)
    print("gauge_R\n", G.R(), "\ngauge_t", G.t(), "\nmax dist", dist.max(), "special", dist_special)


if __name__ == "__main__":
    main()
