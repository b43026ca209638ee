"""Golden vectors of the PCA producer from the LIVE reference
(scedar/eda/sdm.py:1313-1334), generated in the build container:

    python tests/golden/make_golden_pca.py

Cases where sklearn's solver is deterministic LAPACK (``full``: small inputs;
``covariance_eigh``: n >= 10 p, p <= 1000) and one CSR case whose TruncatedSVD
keeps all but one component (the seeded randomized SVD is then exact to 1e-12).
"""
import os
import sys
import warnings

import numpy as np
import scipy.sparse as sps

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader          # noqa: E402

warnings.simplefilter("ignore")
ref = ref_loader.load_reference()
rng = np.random.default_rng(1313)
out = {}


def dense(name, n, p):
    x = rng.standard_normal((n, p)) * np.linspace(3.0, 0.2, p) + \
        rng.standard_normal((1, p))
    out[name + "_x"] = x
    out[name + "_pca"] = ref(x, metric="euclidean")._pca_x


dense("full_60x9", 60, 9)            # full solver, n > p
dense("full_40x120", 40, 120)        # full solver, n < p
dense("eigh_1200x40", 1200, 40)      # covariance_eigh
xs = sps.random(300, 30, density=0.25, format="csr", random_state=7,
                data_rvs=lambda s: rng.lognormal(0.0, 1.0, size=s))
out["csr_300x30_data"] = xs.data
out["csr_300x30_indices"] = xs.indices
out["csr_300x30_indptr"] = xs.indptr
out["csr_300x30_pca"] = ref(xs, metric="cosine")._pca_x
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "pca.npz"), **out)
print({k: v.shape for k, v in out.items()})
