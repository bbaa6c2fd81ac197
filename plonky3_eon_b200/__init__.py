"""plonky3_eon_b200 — B200-native BN254 KZG hot path (batched coset LDE + G1 MSM) behind the
plugin surface of Lolazyx/plonky3-eon (TwoAdicSubgroupDft<Fr>, Pcs<Fr, _> as implemented by KzgPcs).

All compute happens in libeon_kzg.so (hand-written CUDA, sm_100a) through the C ABI declared in
include/eon_kzg.h.  This package is only the host-side mirror of the reference interface.
Importing it does not require a GPU; creating a Context / running anything does.
"""
from .lib import (Context, DegreeTooLarge, EonError, InvalidG1Point, MultiContext, default_context,  # noqa: F401
                  load, pinned_empty)
from .dft import GpuDft  # noqa: F401
from .pcs import GpuKzgPcs, TwoAdicMultiplicativeCoset, observe_commitment  # noqa: F401
from .mmcs import GpuKzgMmcs  # noqa: F401
from . import field  # noqa: F401
