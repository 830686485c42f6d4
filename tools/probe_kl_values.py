"""What is /root/reference/test/data/kl_values.jld?  ("probe it as a pin for row f.3".)

Finding (run in the build container, `python tools/probe_kl_values.py`): the file is a JLD/HDF5 serialisation of ONE
`Celeste.SensitiveFloats.SensitiveFloat{Core.Float64}` with local_P = 32, local_S = 1:
    v  = -31.07655112000045                      (byte offset 8276)
    d  = 32 doubles, 26 non-zero (rows 7..32)    (byte offset 8660)
    h  = 32 x 32, non-zero only in rows/cols 7..32 (byte offset 9252)
Six leading zero rows = position (2) + galaxy shape (4): exactly the parameters the KL term does not touch, so this IS a
KL SensitiveFloat -- but of an OLDER parameterisation: 32 = 2 + 4 + 2 + 2 + 8 + 8 + 2 + k[2 x 2], i.e. a colour prior with
D = 2 mixture components, where the reference at this commit has D = 8 (44 parameters, `param_set.jl:88-107`,
`cfg/{star,gal}_prior.jld` hold 8 components).  No file in the reference reads it (`grep -rn kl_values` finds nothing) and
the D = 2 prior it was computed with is not in the tree, so it cannot pin `elbo_kl.jl` at this commit.  Row f.3 stays
"unpinned" (autograd only); the Julia dumper (tools/julia/dump_golden.jl) writes `elbo_kl_*` records for it instead.
"""
import sys
import numpy as np

path = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/test/data/kl_values.jld"
b = open(path, "rb").read()
print("size", len(b), "header", b[:40])
for key in (b"SensitiveFloat", b"local_P", b"local_S", b"has_hessian"):
    print(key.decode(), "at", b.find(key))
v = np.frombuffer(b, "<f8", 1, 8276)[0]
d = np.frombuffer(b, "<f8", 32, 8708 - 6 * 8)
h = np.frombuffer(b, "<f8", 32 * 32, 9252).reshape(32, 32)
print("v =", v)
print("d: nonzero rows", np.nonzero(d)[0] + 1)
print("h: nonzero rows", np.unique(np.nonzero(h)[0]) + 1, "symmetric", np.allclose(h, h.T))
assert np.count_nonzero(d[:6]) == 0 and np.count_nonzero(d[6:]) == 26
