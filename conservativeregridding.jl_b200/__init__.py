"""B200-native engine behind ConservativeRegridding.jl's Regridder / regrid! hot paths.

Import as ``crg_b200`` (see crg_b200/__init__.py).  Sub-modules:

* ``grids``, ``fields`` -- host-side synthetic grid generators / analytic fields (numpy).
* ``_lib``              -- ctypes binding of the C-ABI library ``libcrgb200.so`` (csrc/).
* ``regridder``         -- Python mirror of the reference API: ``Regridder``, ``regrid_``
                           (= ``regrid!``), ``regrid``, ``transpose``, ``normalize_``.
* ``dist``              -- destination-sharded multi-GPU regridder (torch.distributed).
"""
from . import grids, fields  # noqa: F401
