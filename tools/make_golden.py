"""Generate tests/golden/*.npz: flat ABI-level inputs + oracle outputs for the small parity cases.

    python tools/make_golden.py [case ...]

The generating code path is: tests/cases.py (synthetic inputs, seeded) -> flatten -> oracle
(oracle/libceleste_oracle.so) in modes 0/1/2.  Commit the outputs; tests/test_golden.py checks the
oracle, the emulated kernels and (on the GPU box) the CUDA library against them.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
import golden_io  # noqa: E402
import oracle_lib  # noqa: E402
from celeste_jl_b200.flatten import csr_tasks  # noqa: E402

GOLDEN = ["star_1band", "two_body", "config2", "config2_rotated_wcs", "masked", "clipped_and_empty", "psf_k3",
          "seven_images", "sharp_psf"]

if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    for name in (sys.argv[1:] or GOLDEN):          # python tools/make_golden.py [case ...]
        images, patches, tasks = cases.get(name)
        of = oracle_lib.OracleField(images, patches)
        csr = csr_tasks(tasks)
        outs = {mode: of.elbo_csr(*csr, mode=mode) for mode in (0, 1, 2)}
        path = os.path.join(ROOT, "tests", "golden", name + ".npz")
        golden_io.dump(path, of.fi, of.fp, csr, outs)
        print(name, os.path.getsize(path), "bytes", outs[2]["v"])
