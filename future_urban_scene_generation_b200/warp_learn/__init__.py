"""Drop-in mirror of the reference's `warp_learn` package for the hot path:
`online_visibility` and `planes_utils` with the reference's names and signatures, plus the batched
entry point `warp_batch` that the fused sm_100a kernels were built for."""
from .batch import warp_batch, check_refused, RefusedCrops, WarpResult  # noqa: F401
