"""Generates the golden vectors under tests/golden/ (run once in the build container; committed).

The reference (/root/reference) cannot be compiled or imported here, so the fixtures are the
known-answer inputs written in its test sources plus expected values computed by an independent
trusted implementation (scipy / numpy), exactly what the reference's own tests compare against
(Eigen's MatrixBase::exp / EigenSolver):
  * expm_4x4.json    — conex/test/exponential_map_pade_test.cc:19-27
  * lanczos_4x4.json — conex/test/approximate_eigenvalues.cc:18-32,65-85
"""
import json
import os

import numpy as np
import scipy.linalg

here = os.path.dirname(os.path.abspath(__file__))
A = np.array([[3, 1, 0, 1], [1, 3, 1, 0], [0, 1, 4, 1], [1, 0, 1, 5]], float)
A /= np.trace(A)
with open(os.path.join(here, "expm_4x4.json"), "w") as f:
    json.dump({"A": A.tolist(), "expm": scipy.linalg.expm(A).tolist()}, f, indent=1)

rng = np.random.Generator(np.random.PCG64(0))
R = rng.uniform(-1, 1, (4, 4))
W = R @ R.T
with open(os.path.join(here, "lanczos_4x4.json"), "w") as f:
    json.dump({"A": A.tolist(), "W": W.tolist(), "r0": [1.0, 2.0, 0.0, 4.0],
               "eig_WA": np.sort(np.linalg.eigvals(W @ A).real).tolist(),
               "eig_A": np.sort(np.linalg.eigvalsh(A)).tolist()}, f, indent=1)
print("wrote golden vectors")
