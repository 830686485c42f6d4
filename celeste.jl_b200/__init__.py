"""celeste.jl_b200 -- B200-native (sm_100a) implementation of Celeste.jl's per-source
ELBO hot path behind the reference's own API surface (ElboArgs / elbo_likelihood / elbo).

The directory name carries a dot, so import it through the `celeste_jl_b200` shim at the
repository root (`import celeste_jl_b200 as cj`).  Host-side modules mirror the reference:
  model            <- src/model/*.jl                (Image, ImagePatch, PsfComponent, ids, ...)
  deterministic_vi <- src/DeterministicVI.jl, src/deterministic_vi/elbo_*.jl
  synthetic        <- src/Synthetic.jl, test/SampleData.jl
  parallel_run     <- src/partition.jl, src/ParallelRun.jl:28-95 (partitioning / multi-GPU sharding)
  csrc/            -- the CUDA kernels and the C ABI (include/celeste_cuda.h)
"""
from . import _lib, model, flatten, deterministic_vi, parallel_run, constraint_transforms, kl, elbo_maximize  # noqa: F401
from .model import (AffineWCS, CatalogEntry, Image, ImagePatch, PsfComponent, ids, get_sky_patches,  # noqa: F401
                    find_neighbors, find_all_neighbors)
from .deterministic_vi import (DeviceField, ElboArgs, ElboIntermediateVariables, Plan, SensitiveFloat,  # noqa: F401
                               catalog_init_source, elbo, elbo_likelihood, fill_celeste_expectation,
                               generic_init_source, init_sources)

__all__ = ["model", "deterministic_vi", "flatten", "ElboArgs", "elbo", "elbo_likelihood", "SensitiveFloat",
           "DeviceField", "Plan", "Image", "ImagePatch", "PsfComponent", "CatalogEntry", "ids"]
