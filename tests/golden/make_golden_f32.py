"""Golden for the notebook's float32 call (main.ipynb cell 7 passes ``dtype=np.float32``): the REAL reference with
``dtype=np.float32`` on the ``net_small_cg_it3`` inputs.  ``python tests/golden/make_golden_f32.py`` ->
``f32_net_small_cg_it3.npz`` (outputs only; the inputs are those of ``net_small_cg_it3.npz``)."""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..")))
sys.path.insert(0, "/root/reference")
os.environ["TQDM_DISABLE"] = "1"

from vican.bipgo import bipartite_se3sync        # noqa: E402  (reference)
from vican.geometry import SE3 as RefSE3          # noqa: E402  (reference)

from vican_b200 import synthetic as syn           # noqa: E402
from util import load_golden                      # noqa: E402


def main():
    g, params, filter_on, _ = load_golden("net_small_cg_it3")
    edges, constraints = syn.to_edge_dict(g, RefSE3)
    nr, nt, ef = syn.default_callables()
    last = None
    for _ in range(8):
        try:
            with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
                out = bipartite_se3sync(edges, constraints=constraints, noise_model_r=nr, noise_model_t=nt,
                                        edge_filter=ef, dtype=np.float32, **params)
            break
        except np.linalg.LinAlgError as exc:
            last = exc
    else:
        raise last
    keys = sorted(out.keys())
    k0 = keys[0]
    print("dtypes of the reference output: R", out[k0].R().dtype, " t", out[k0].t().dtype)
    np.savez_compressed(os.path.join(HERE, "f32_net_small_cg_it3.npz"), out_keys=np.array(keys),
                        out_R=np.stack([np.asarray(out[k].R(), np.float64) for k in keys]),
                        out_t=np.stack([np.asarray(out[k].t(), np.float64) for k in keys]),
                        R_dtype=str(out[k0].R().dtype), t_dtype=str(out[k0].t().dtype))


if __name__ == "__main__":
    main()
