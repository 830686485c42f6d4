"""Extract the colour-prior constants from the reference's cfg/*.jld (HDF5) files.

Run once in the build container (needs /root/reference); output is committed as
celeste.jl_b200/data/celeste_priors.json.  The raw little-endian float64 payloads
sit at fixed byte offsets in both files: c_covs 4x4x8 @4220,
c_means 4x8 @5612, c_weights 8 @6540 (Julia column-major).  Consumed by
light_source_model.jl:90-133 `load_prior_init` in the reference.
"""
import json
import sys
import numpy as np

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
out = {}
for name, fn in (("star", "star_prior.jld"), ("gal", "gal_prior.jld")):
    b = open(f"{ref}/cfg/{fn}", "rb").read()
    covs = np.frombuffer(b, dtype="<f8", count=128, offset=4220).reshape(8, 4, 4)  # [k][c2][c1]
    means = np.frombuffer(b, dtype="<f8", count=32, offset=5612).reshape(8, 4)     # [k][c]
    w = np.frombuffer(b, dtype="<f8", count=8, offset=6540)
    assert abs(w.sum() - 1.0) < 1e-12
    for k in range(8):
        assert np.allclose(covs[k], covs[k].T) and np.linalg.eigvalsh(covs[k]).min() > 0
    out[name] = {"c_weights": w.tolist(), "c_means": means.tolist(), "c_covs": covs.tolist()}
json.dump(out, open("celeste.jl_b200/data/celeste_priors.json", "w"), indent=0)
print("ok")
