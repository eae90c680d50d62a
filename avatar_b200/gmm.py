"""GaussianMixture: host-side mirror of ark::GaussianMixture's file format (GaussianMixture.cpp:12-58).

Only the parsing lives here; the Cholesky / normalisation maths of load() runs inside the library
(avb_model_create), because that is what the kernels consume.
"""
import numpy as np


class GaussianMixture:
    def __init__(self):
        self.nComps = -1   # GaussianMixture.cpp:15-19: missing file => nComps = -1
        self.nDims = 0
        self.weight = None
        self.mean = None
        self.cov = None

    @classmethod
    def from_arrays(cls, weight, mean, cov):
        g = cls()
        g.weight = np.ascontiguousarray(weight, dtype=np.float64)
        g.mean = np.ascontiguousarray(mean, dtype=np.float64)
        g.cov = np.ascontiguousarray(cov, dtype=np.float64)
        g.nComps, g.nDims = g.mean.shape
        assert g.cov.shape == (g.nComps, g.nDims, g.nDims)
        return g

    def load(self, path):
        """text format: `C D`, C weights, C*D means, C*D*D covariances (GaussianMixture.cpp:20-58)"""
        try:
            with open(path) as fh:
                tok = fh.read().split()
        except OSError:
            print(f"Warning: pose prior file at {path} does not exist or cannot be read")
            self.nComps = -1
            return
        C, D = int(tok[0]), int(tok[1])
        vals = np.array(tok[2:2 + C + C * D + C * D * D], dtype=np.float64)
        self.nComps, self.nDims = C, D
        self.weight = vals[:C].copy()
        self.mean = vals[C:C + C * D].reshape(C, D).copy()
        self.cov = vals[C + C * D:].reshape(C, D, D).copy()

    def save(self, path):
        with open(path, "w") as fh:
            fh.write(f"{self.nComps} {self.nDims}\n")
            fh.write(" ".join(repr(float(x)) for x in self.weight) + "\n")
            for row in self.mean:
                fh.write(" ".join(repr(float(x)) for x in row) + "\n")
            for c in self.cov:
                for row in c:
                    fh.write(" ".join(repr(float(x)) for x in row) + "\n")

    def numComponents(self):
        return self.nComps
